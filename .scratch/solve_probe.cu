#include "../kiss-icp_b200/csrc/device_ops.cuh"
#include "../kiss-icp_b200/csrc/icp_team.cuh"
using namespace kb;
__global__ void k_probe(const double *in, double *out, double conv) {
    __shared__ Shared sh;
    for (int i = 0; i < NPART; ++i) sh.red[i] = in[i];
    team_solve(sh, conv, false);
    out[0] = sh.pending.q.x; out[1] = sh.pending.q.y; out[2]=sh.pending.q.z; out[3]=sh.pending.q.w;
    out[4] = sh.pending.t.x; out[5]=sh.pending.t.y; out[6]=sh.pending.t.z; out[7]=sh.flag;
}
