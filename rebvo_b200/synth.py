"""Seeded synthetic inputs for the REBVO edge pipeline (no dataset is available offline).

SURVEY.md section 8(d): G(seed) = axis-aligned filled rectangles + N(0,1.5) noise, r=g=b.  Sequences
are rendered from two depth layers (a far background canvas and a nearer foreground layer of opaque
rectangles) seen by a camera translating on a smooth path parallel to the image plane, so that the
stream has real parallax for the inverse-depth EKF.  Pure numpy; used by tests/ and bench.py.
"""
import numpy as np

EUROC = dict(w=752, h=480, zfx=458.654, zfy=457.296, ppx=367.215, ppy=248.375)
TUM = dict(w=640, h=480, zfx=525.0, zfy=525.0, ppx=320.0, ppy=240.0)


def rect_canvas(rng, w, h, nrect, base=None, alpha=False, smin=8, smax=90):
    """Paint `nrect` filled rectangles (centre ~U, sides ~U(smin,smax), grey ~U(20,235)) in order."""
    img = np.full((h, w), 128.0 if base is None else base, np.float32)
    a = np.zeros((h, w), np.float32)
    for _ in range(nrect):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        sx, sy = rng.uniform(smin, smax), rng.uniform(smin, smax)
        g = rng.uniform(20, 235)
        x0, x1 = int(max(0, cx - sx / 2)), int(min(w, cx + sx / 2))
        y0, y1 = int(max(0, cy - sy / 2)), int(min(h, cy + sy / 2))
        img[y0:y1, x0:x1] = g
        a[y0:y1, x0:x1] = 1.0
    return (img, a) if alpha else img


def to_rgb_u8(gray_f, rng=None, noise=1.5):
    g = gray_f
    if rng is not None and noise > 0:
        g = g + rng.normal(0.0, noise, g.shape).astype(np.float32)
    g8 = np.clip(np.rint(g), 0, 255).astype(np.uint8)
    return np.repeat(g8[:, :, None], 3, axis=2).copy()


def frame_pair(seed=42, w=640, h=480, nrect=220, shift=(1.5, 0.7)):
    """Config 1 stand-in: frame 2 = same scene shifted by `shift` px (bilinear), fresh noise each."""
    rng = np.random.default_rng(seed)
    m = 8
    canvas = rect_canvas(rng, w + 2 * m, h + 2 * m, nrect)
    f1 = canvas[m:m + h, m:m + w]
    f2 = _crop_bilinear(canvas, m - shift[0], m - shift[1], w, h)
    return to_rgb_u8(f1, rng), to_rgb_u8(f2, rng)


def _crop_bilinear(canvas, x0, y0, w, h):
    xi, yi = int(np.floor(x0)), int(np.floor(y0))
    a, b = np.float32(x0 - xi), np.float32(y0 - yi)
    c = canvas[yi:yi + h + 1, xi:xi + w + 1]
    return ((1 - a) * (1 - b) * c[:h, :w] + a * (1 - b) * c[:h, 1:w + 1]
            + (1 - a) * b * c[1:h + 1, :w] + a * b * c[1:h + 1, 1:w + 1])


class Sequence:
    """A seeded two-layer parallax stream.  frame(i) -> (timestamp, HxWx3 uint8)."""

    def __init__(self, w=752, h=480, seed=7, zf=458.0, fps=20.0, nrect_bg=420, nrect_fg=60,
                 z_bg=2.0, z_fg=1.0, speed=0.006, noise=1.5):
        self.w, self.h, self.seed, self.zf, self.fps, self.noise = w, h, seed, zf, fps, noise
        self.z_bg, self.z_fg, self.speed = z_bg, z_fg, speed
        rng = np.random.default_rng(seed)
        self.m = m = 200
        self.bg = rect_canvas(rng, w + 2 * m, h + 2 * m, nrect_bg)
        self.fg, self.fa = rect_canvas(rng, w + 2 * m, h + 2 * m, nrect_fg, alpha=True, smin=30, smax=120)

    def cam_pos(self, i):
        """Camera translation (metres) at frame i: a smooth Lissajous path in the image plane."""
        t = i / self.fps
        s = self.speed * self.fps
        return np.array([0.12 * np.sin(s / 0.12 * t * 0.35), 0.07 * np.sin(s / 0.07 * t * 0.22 + 0.5), 0.0])

    def frame(self, i):
        p = self.cam_pos(i)
        m, w, h = self.m, self.w, self.h
        sb = self.zf * p[:2] / self.z_bg
        sf = self.zf * p[:2] / self.z_fg
        bg = _crop_bilinear(self.bg, m + sb[0], m + sb[1], w, h)
        fg = _crop_bilinear(self.fg, m + sf[0], m + sf[1], w, h)
        fa = _crop_bilinear(self.fa, m + sf[0], m + sf[1], w, h)
        g = bg * (1 - fa) + fg * fa
        rng = np.random.default_rng((self.seed + 1) * 1000003 + i)
        return i / self.fps, to_rgb_u8(g, rng, self.noise)

    def frames(self, n, start=0):
        ts = np.empty(n, np.float64)
        out = np.empty((n, self.h, self.w, 3), np.uint8)
        for k in range(n):
            ts[k], out[k] = self.frame(start + k)
        return ts, out


def write_frames_file(path, ts, frames):
    """Raw frame file read by oracle/ref_driver.cpp: int32 W,H,N then per frame f64 t + RGB24."""
    n, h, w, _ = frames.shape
    with open(path, "wb") as f:
        f.write(np.array([w, h, n], np.int32).tobytes())
        for i in range(n):
            f.write(np.float64(ts[i]).tobytes())
            f.write(frames[i].tobytes())


def imu_samples(seq, n_frames, rate=200.0, gyro_bias=(0.02, -0.015, 0.01), gyro_noise=1.7e-4, acc_noise=2e-3, seed=0,
                g=9.8):
    """Synthetic IMU stream for a Sequence (SURVEY.md 8(d) config 3 fallback): samples at `rate` Hz covering the
    frames, columns t[s], gyro xyz [rad/s], accel xyz [m/s^2] in the camera frame.  The synthetic camera does not rotate:
    the gyroscope measures bias + noise; the accelerometer measures the second derivative of the camera path minus
    gravity (0, g, 0) (the filter's model a_s + g = k * a_v, scaleestimator.cpp:119-121)."""
    rng = np.random.default_rng(seed + 77)
    t = np.arange(-0.25, n_frames / seq.fps + 0.25, 1.0 / rate)
    h = 1e-3

    def pos(tt):
        return np.stack([seq.cam_pos(x * seq.fps) for x in tt])

    acc = (pos(t + h) - 2 * pos(t) + pos(t - h)) / (h * h)
    acc = acc - np.array([0.0, g, 0.0])
    acc = acc + rng.normal(0, acc_noise, acc.shape)
    gyro = np.array(gyro_bias)[None, :] + rng.normal(0, gyro_noise, (len(t), 3))
    return np.concatenate([t[:, None], gyro, acc], axis=1)


def write_imu_csv(path, samples):
    """csv read by ImuGrabber::LoadDataSet (src/UtilLib/imugrabber.cpp:80-132): t,gx,gy,gz,ax,ay,az per line."""
    with open(path, "w") as f:
        f.write("#timestamp,w_x,w_y,w_z,a_x,a_y,a_z\n")
        for r in samples:
            f.write(",".join("%.17g" % v for v in r) + "\n")


def imu_samples_walk(seq, total, base_n, rate=200.0, gyro_bias=(0.02, -0.015, 0.01), gyro_noise=1.7e-4, acc_noise=2e-3, seed=0,
                     g=9.8):
    """IMU stream for a sequence whose `total` frames walk `base_n` rendered frames back and forth (bench.py): the camera
    position at time t is cam_pos at the triangle-wave frame index."""
    rng = np.random.default_rng(seed + 78)
    period = max(1, 2 * (base_n - 1))
    t = np.arange(-0.25, total / seq.fps + 0.25, 1.0 / rate)
    h = 1e-3

    def pos(tt):
        x = (tt * seq.fps) % period
        x = np.where(x < base_n - 1, x, period - x)
        return np.stack([seq.cam_pos(v) for v in x])

    acc = (pos(t + h) - 2 * pos(t) + pos(t - h)) / (h * h)
    acc = np.clip(acc, -20, 20) - np.array([0.0, g, 0.0])
    acc = acc + rng.normal(0, acc_noise, acc.shape)
    gyro = np.array(gyro_bias)[None, :] + rng.normal(0, gyro_noise, (len(t), 3))
    return np.concatenate([t[:, None], gyro, acc], axis=1)
