// TEST-ONLY: runs the team-per-problem DLS iteration (ik_b200/csrc/dls_team.cuh -- the very source the CUDA kernel
// dls_team_kernel instantiates) on the CPU: 16 host threads play the 16 lanes of a team, a std::barrier plays
// __syncwarp.  Never linked into libikb200.so; the product has no CPU path.
#include <barrier>
#include <thread>
#include <vector>

#include "../../ik_b200/csrc/dls_team.cuh"

using namespace ikb;

// consts: PR[2][7][9], Pp[2][7][3], FR[2][9], Fp[2][3] flattened in that order (186 doubles)
template <typename T>
static int team_solve(const double *tree, const double *lower, const double *upper, const double *weight, const double *q0,
                      const double *targets, int max_it, double step, double damping, double tol, double *q_out, int *iters,
                      double *resid, double *e_first, double *J_first) {
    TeamConsts<T> c{};
    const double *p = tree;
    for (int l = 0; l < 2; ++l) for (int k = 0; k < kTeamChain; ++k) for (int i = 0; i < 9; ++i) c.PR[l][k][i] = (T)*p++;
    for (int l = 0; l < 2; ++l) for (int k = 0; k < kTeamChain; ++k) for (int i = 0; i < 3; ++i) c.Pp[l][k][i] = (T)*p++;
    for (int l = 0; l < 2; ++l) for (int i = 0; i < 9; ++i) c.FR[l][i] = (T)*p++;
    for (int l = 0; l < 2; ++l) for (int i = 0; i < 3; ++i) c.Fp[l][i] = (T)*p++;
    for (int k = 0; k < 23; ++k) { c.lower[k] = (T)lower[k]; c.upper[k] = (T)upper[k]; }
    for (int i = 0; i < 12; ++i) c.weight[i] = (T)weight[i];
    TeamScratch<T> S{};
    for (int i = 0; i < 36; ++i) S.tg[i] = (T)targets[i];
    std::barrier<> bar(kTeamLanes);
    int it_out = 0, success = 0;
    T res_out = 0;
    std::vector<std::thread> th;
    for (int lane = 0; lane < kTeamLanes; ++lane)
        th.emplace_back([&, lane]() {
            TeamLane<T> st;
            for (int k = 0; k < 7; ++k) st.qff[k] = (T)q0[k];
            st.qr = (T)q0[7 + lane];
            st.lo = c.lower[7 + lane];
            st.hi = c.upper[7 + lane];
            st.wgt = c.weight[lane < 12 ? lane : 0];
            auto sync = [&]() { bar.arrive_and_wait(); };
            int it = 0, ok = 0;
            T res = 0;
            while (it < max_it) {
                res = team_iteration(lane, st, S, c, (T)step, (T)(damping * damping), (T)tol, sync);
                if (it == 0 && lane == 0) {
                    if (e_first) for (int i = 0; i < 12; ++i) e_first[i] = (double)S.e[i];
                    if (J_first) for (int a = 0; a < 12; ++a) for (int k = 0; k < 13; ++k) J_first[13 * a + k] = (double)S.J[a][k];
                }
                if (res < (T)tol) { ok = 1; break; }
                ++it;
            }
            bar.arrive_and_wait();
            for (int k = 0; k < 7; ++k) if (lane == k) q_out[k] = (double)st.qff[k];
            q_out[7 + lane] = (double)st.qr;
            if (lane == 0) { it_out = it; success = ok; res_out = res; }
        });
    for (auto &t : th) t.join();
    *iters = it_out;
    *resid = (double)res_out;
    return success;
}

extern "C" int h_team_cassie_d(const double *tree, const double *lo, const double *hi, const double *w, const double *q0,
                               const double *tg, int mi, double st, double da, double tol, double *q, int *it, double *res,
                               double *e0, double *J0) {
    return team_solve<double>(tree, lo, hi, w, q0, tg, mi, st, da, tol, q, it, res, e0, J0);
}
extern "C" int h_team_cassie_f(const double *tree, const double *lo, const double *hi, const double *w, const double *q0,
                               const double *tg, int mi, double st, double da, double tol, double *q, int *it, double *res,
                               double *e0, double *J0) {
    return team_solve<float>(tree, lo, hi, w, q0, tg, mi, st, da, tol, q, it, res, e0, J0);
}
