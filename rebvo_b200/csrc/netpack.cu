// netpack.cu -- wire egress of an edge map: the 15-byte packed net_keyline records the
// reference's third thread builds by walking all 168-byte KeyLine structs on the host every frame
//   copy_net_keyline        src/CommLib/net_keypoint.cpp:29-75   (from_pair == nullptr: monocular)
//   copy_net_keyline_nextid src/CommLib/net_keypoint.cpp:79-107
//   struct net_keyline      include/CommLib/net_keypoint.h:37-62 (#pragma pack(1): qx qy rho s_rho | n_kl | m_num | flow.x flow.y)
// packed here straight from the device SoA: a consumer that only forwards edge maps moves 15 bytes per keyline over
// PCIe instead of 168.  Integer / byte work: bit-exact (tests/test_gpu_netpack.py against the reference's packer).
#include "common.cuh"

#define NET_RHO_SCALING 10000.0   // net_keypoint.h:33

__device__ __forceinline__ unsigned short d_clamp_ushort(float f) {   // util.h:58-66
    if (f < 0) return 0;
    if (f > 65535.0) return 65535;
    return (unsigned short)f;
}
__device__ __forceinline__ unsigned char d_clamp_uchar(float f) {     // util.h:52-55: (f<0?0:f>255.0?255:f) evaluates in float
    const float r = f < 0 ? 0.f : (f > 255.0 ? 255.f : f);
    return (unsigned char)r;
}

__device__ __forceinline__ void pack_net_one(const KLSoA &kl, int j, int n, double k_prof, unsigned char *__restrict__ out) {
    const float2 cp = kl.c_p[j], pm = kl.p_m[j], pm0 = kl.p_m_0[j];
    const unsigned short qx = (unsigned short)round((double)cp.x), qy = (unsigned short)round((double)cp.y);
    unsigned short rho = d_clamp_ushort((float)(NET_RHO_SCALING * kl.rho[j] / k_prof));
    unsigned short s_rho = d_clamp_ushort((float)(NET_RHO_SCALING * kl.s_rho[j] / k_prof));
    rho = rho > 1 ? rho : 1;        // std::max(..., (u_short)1)
    s_rho = s_rho > 1 ? s_rho : 1;
    const unsigned char fx = d_clamp_uchar((float)round((double)((pm.x - pm0.x) * 10) + 127.0));
    const unsigned char fy = d_clamp_uchar((float)round((double)((pm.y - pm0.y) * 10) + 127.0));
    // copy_net_keyline numbers the records in keyline order (net_id = j), so copy_net_keyline_nextid's
    // to[j].n_kl = from[kl.n_id].net_id is n_id itself where that record exists; it stays -1 otherwise
    const int nid = kl.n_id[j];
    const int n_kl = (nid >= 0 && nid < n) ? nid : -1;
    const unsigned char m_num = d_clamp_uchar((float)kl.m_num[j]);
    unsigned char *o = out + (size_t)j * 15;
    o[0] = (unsigned char)(qx & 0xff);
    o[1] = (unsigned char)(qx >> 8);
    o[2] = (unsigned char)(qy & 0xff);
    o[3] = (unsigned char)(qy >> 8);
    o[4] = (unsigned char)(rho & 0xff);
    o[5] = (unsigned char)(rho >> 8);
    o[6] = (unsigned char)(s_rho & 0xff);
    o[7] = (unsigned char)(s_rho >> 8);
    o[8] = (unsigned char)(n_kl & 0xff);
    o[9] = (unsigned char)((n_kl >> 8) & 0xff);
    o[10] = (unsigned char)((n_kl >> 16) & 0xff);
    o[11] = (unsigned char)((n_kl >> 24) & 0xff);
    o[12] = m_num;
    o[13] = fx;
    o[14] = fy;
}

__global__ void __launch_bounds__(256) k_pack_net(KLSoA kl, const MapState *__restrict__ st, unsigned char *__restrict__ out,
                                                  int capacity, double k_prof, int *n_out) {
    const int kn = st->kn;
    const int n = kn < capacity ? kn : capacity;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) *n_out = n;
    if (j >= n) return;
    pack_net_one(kl, j, n, k_prof, out);
}

// pipeline mirror (rb_pipeline_set_mirror mode 2): destination through a device pointer, K from the frame's nav record
__global__ void __launch_bounds__(256) k_pack_net_ind(KLSoA kl, const MapState *__restrict__ st, unsigned char *const *base,
                                                      size_t offset, const double *__restrict__ k_prof) {
    const int n = st->kn;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    pack_net_one(kl, j, n, *k_prof, *base + offset);
}
int rb_map_pack_net_enqueue(rb_ctx *c, rb_map *m, unsigned char *const *base, size_t offset, const double *k_prof) {
    k_pack_net_ind<<<rb_div_up(c->kcap, 256), 256, 0, c->stream>>>(m->kl, m->st, base, offset, k_prof);
    RB_LAUNCH_CHECK();
    return RB_OK;
}

extern "C" int rb_map_pack_net_keylines(rb_map *m, double k_prof, void *dst, int capacity, int *n_out) {
    if (!m || !dst || capacity < 0) return RB_ERR_ARG;
    RB_ENTER(m->c);
    rb_ctx *c = m->c;
    unsigned char *tmp = nullptr;
    int *nd = (int *)((char *)c->dev_small + RB_DS_ARGS + 2048);
    const int cap = capacity < c->kcap ? capacity : c->kcap;
    if (cap > 0) RB_CUDA(cudaMalloc(&tmp, (size_t)cap * 15));
    k_pack_net<<<rb_div_up(cap > 0 ? cap : 1, 256), 256, 0, c->stream>>>(m->kl, m->st, tmp, cap, k_prof, nd);
    c->launches++;
    int n = 0;
    cudaError_t e = cudaMemcpyAsync(&n, nd, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess && n > 0) e = cudaMemcpy(dst, tmp, (size_t)n * 15, cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (e != cudaSuccess) {
        snprintf(c->err, sizeof(c->err), "pack_net_keylines: %s", cudaGetErrorString(e));
        return RB_ERR_CUDA;
    }
    if (n_out) *n_out = n;
    return RB_OK;
}
