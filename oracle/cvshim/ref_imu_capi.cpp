// oracle/cvshim/ref_imu_capi.cpp — C entry around the reference's own Imu class (TEST INFRASTRUCTURE).  Built by
// `make -C oracle ref` with /root/reference/src/Imu.cpp and Plus.cpp (unmodified) against the cv and ROS shims into
// oracle/_ref/libref_imu.so.  The external Madgwick node is replaced by a table of orientations supplied by the caller
// (one quaternion per published sample, in publication order): what is pinned here is everything Imu.cpp itself does.
#include "Imu.hpp"
#include <iostream>
#include <sstream>

namespace {
struct Quiet {
    std::streambuf* old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf()) { std::cout.rdbuf(sink.rdbuf()); }
    ~Quiet() { std::cout.rdbuf(old); }
};
struct Table { const double* q_wxyz; int n, next; };
void hook(void* user, const sensor_msgs::Imu&, sensor_msgs::Imu& fused) {
    Table* t = static_cast<Table*>(user);
    if (t->next < t->n) {
        const double* q = t->q_wxyz + 4 * t->next++;
        fused.orientation.w = q[0]; fused.orientation.x = q[1]; fused.orientation.y = q[2]; fused.orientation.z = q[3];
    }
}
void pts(const double* v, int n, std::vector<Point3d>& out) {
    out.clear();
    for (int i = 0; i < n; i++) out.push_back(Point3d(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
}
}  // namespace

// One Imu life: initializate(gt_yaw, gt_velocity, first n0 samples) then `steps` x estimate() on consecutive blocks of
// n_per samples.  w, a: (n0 + steps * n_per) x 3; q_wxyz: one orientation per sample (what the filter node answered).
// out, per step (and once after initializate, step index 0): 60 doubles
//   [0..2] residualRPY, [3..5] residualPosition, [6..8] residualVelocity, [9..11] velocity, [12..14] position,
//   [15..17] angBias, [18..20] accBias, [21..23] initialVelocity, [24..32] init R, [33..41] final R, [42..50] residual R,
//   [51..53] rpyAnglesWorld.back(), [54..56] accelerationWorld.back(), [57..59] angularVelocity
extern "C" int ref_imu_run(double timestep, double gt_yaw, const double gt_vel[3], const double* w, const double* a,
                           const double* q_wxyz, int n0, int n_per, int steps, double* out) {
    Quiet quiet;
    rosshim::bus() = rosshim::Bus();
    Table table = {q_wxyz, n0 + steps * n_per, 0};
    rosshim::bus().hook = hook;
    rosshim::bus().user = &table;
    Imu imu(timestep);
    std::vector<Point3d> wv, av;
    pts(w, n0, wv); pts(a, n0, av);
    imu.initializate(gt_yaw, Point3d(gt_vel[0], gt_vel[1], gt_vel[2]), wv, av);
    auto dump = [&](double* o) {
        const Point3d* p[] = {&imu.residualRPY, &imu.residualPosition, &imu.residualVelocity, &imu.velocity, &imu.position,
                              &imu.angBias, &imu.accBias, &imu.initialVelocity};
        for (int k = 0; k < 8; k++) { o[3 * k] = p[k]->x; o[3 * k + 1] = p[k]->y; o[3 * k + 2] = p[k]->z; }
        for (int i = 0; i < 9; i++) {
            o[24 + i] = imu.init_rotationMatrix.val[i];
            o[33 + i] = imu.final_rotationMatrix.val[i];
            o[42 + i] = imu.residual_rotationMatrix.val[i];
        }
        const Point3d r = imu.rpyAnglesWorld.empty() ? Point3d() : imu.rpyAnglesWorld.back();
        const Point3d aw = imu.accelerationWorld.empty() ? Point3d() : imu.accelerationWorld.back();
        o[51] = r.x; o[52] = r.y; o[53] = r.z;
        o[54] = aw.x; o[55] = aw.y; o[56] = aw.z;
        o[57] = imu.angularVelocity.x; o[58] = imu.angularVelocity.y; o[59] = imu.angularVelocity.z;
    };
    dump(out);
    for (int s = 0; s < steps; s++) {
        pts(w + 3 * (n0 + s * n_per), n_per, wv);
        pts(a + 3 * (n0 + s * n_per), n_per, av);
        imu.setImuData(wv, av);
        imu.estimate();
        dump(out + 60 * (s + 1));
    }
    rosshim::bus() = rosshim::Bus();
    return 0;
}
