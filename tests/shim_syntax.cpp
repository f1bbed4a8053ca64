// compile-only: instantiate the IMU-mode mirrors and the copy constructors of the shim
#include <rebvo_b200_shim.hpp>
using namespace rebvo;
void instantiate(global_tracker &gt, edge_tracker &a) {
    TooN::Vector<3> V = TooN::Zeros;
    TooN::Matrix<3, 3> R = TooN::Identity, Wb = TooN::Identity;
    TooN::Matrix<6, 6> Wx = TooN::Identity, Rx = TooN::Identity;
    TooN::Vector<6> X = TooN::Zeros;
    TooN::Vector<3> Gb = TooN::Zeros;
    gt.Minimizer_V<double>(V, R, a, 0.5, 5, 1.0, 2u, 1.0, 0.f);
    gt.Minimizer_V<float>(V, R, a, 0.5f, 5, 1.0f, 2u, 1.0, 0.f);
    a.ExtRotVel(V, Wx, Rx, X, 1.0, 1.0);
    edge_tracker::BiasCorrect(X, Wx, Gb, Wb, R, R);
    // keyframe.cpp:28-35: deep copies of the edge map and its tracker
    edge_tracker *copy = new edge_tracker(a);
    global_tracker *gcopy = new global_tracker(gt);
    gcopy->SetEdgeTracker(copy);
    delete gcopy;
    delete copy;
}
