// Team-per-problem DLS IK iteration: the LATENCY path of the specialised solver.
//
// The thread-per-problem kernel (dls_spec.cuh) is a throughput design: one DLS iteration is ~3 500 dependent
// instructions of ONE thread, 7 us however few problems are left.  The reference lets a non-converging problem run all
// max_iterations = 100 steps (dls.cpp:14), so the last few per cent of a batch -- and every batch too small to fill the
// GPU -- are bound by 100 x that latency, not by arithmetic.  Here a TEAM of 16 lanes (half a warp) shares one problem and
// every phase of ik::dls is split by DATA over the lanes (same instruction stream, SPMD), so the dependent chain per
// iteration shrinks ~3x:
//
//   phase 0  every lane: sin/cos of "its" revolute joint (16 lanes <-> Cassie's 16 revolutes) -> scratch; base rotation
//            from the quaternion (replicated, every lane carries the 7 free-flyer coordinates).
//   phase 1  forward kinematics by ROWS: row i of R_world(joint) and component i of p_world(joint) depend only on row i
//            of the parent's rotation (R_j = R_parent * const * Rz(q_j)), so lane (limb, i) walks the limb chain with 3 + 1
//            scalars and no communication.  It publishes the joint axes / origins (for the Jacobian) and row i of the task
//            frame.                                                                         (data.cpp:28-30)
//   phase 2  lane a <-> task row a (pelvis linear 0-2, pelvis angular 3-5, left foot 6-8, right foot 9-11): SE3 log error
//            (frame.hpp:37-62) and the Jlog6 blocks of its task (replicated per task row -- scalar chains, nothing to
//            split), row a of M1 = A Rf^T and M2 = B Rf^T.
//   phase 3  row a of the weighted task Jacobian, -(M1 ((p_j - p_f) x z_j) + M2 z_j) per supporting column (frame.hpp:
//            152-182 without materialising the 6 x nv frame Jacobian), kept in registers and published.
//   phase 4  row a of the Gram matrix J J^T + damping^2 I (dls.cpp:39-41); lane 12 takes the right-hand side e as row 12.
//   phase 5  LDL^T by rows, right-looking: per pivot k every lane publishes its column-k entry, one team barrier, then
//            eliminates its own row.  Lane 12's multipliers are D^-1 L^-1 e.                  (dls.cpp:53, Eigen ldlt())
//   phase 6  multipliers -> scratch; every lane back-substitutes y = L^-T (.) redundantly (66 FMAs, no barriers).
//   phase 7  dq = -J^T y (dls.cpp:52): lane l its revolute column, lanes 0-5 the free-flyer columns (published).
//   phase 8  integrate + clamp (dls.cpp:67-71): revolutes per lane, the free-flyer exp6 update replicated.
//
// Lanes exchange data through a 4.7 KB block of shared memory per team; the barrier is __syncwarp (both teams of a warp
// run the same phases in lock step).  17 barriers per iteration.
//
// The body is __host__ __device__ and takes the barrier as a functor, so tests/cpu_harness runs the very same source with
// 16 host threads and a std::barrier per team and checks it against the oracle on the GPU-less build box.
//
// Scope: a free-flyer base with a Full frame task on the base frame and two limbs that are serial chains of seven
// z-axis revolutes ending in a Position frame task -- Cassie with the BASELINE.json task set (spec_cassie_feet_pelvis.cu
// fills the constants from the generated specialisation's signature, so it is exactly the problem spec_matches accepted).
#pragma once
#include "dev_problem.hpp"
#include "se3_math.cuh"

namespace ikb {

constexpr int kTeamLanes = 16;
constexpr int kTeamChain = 7;  // revolute joints per limb chain

// Model constants of the team path (passed by value as a kernel parameter, staged into shared memory once per CTA).
template <typename T> struct alignas(16) TeamConsts {
    T PR[2][kTeamChain][9];  // placement rotation of chain joint k of limb l (row-major)
    T Pp[2][kTeamChain][3];  //   ... translation
    T FR[2][9], Fp[2][3];    // task frame placement on the last chain joint
    T lower[24], upper[24];  // position limits (nq = 23)
    T weight[12];            // task row weights, stacked order
};

// Per-team exchange block (shared memory on the device).  Every array starts 16-byte aligned.
template <typename T> struct alignas(16) TeamScratch {
    T sc[16][2];      // (sin, cos) of revolute joint l
    T frame[3][12];   // world placement of the three task frames: R row-major (9) + p (3)
    T z[2][kTeamChain][4];  // world axis of chain joint k of limb l (+ pad)
    T p[2][kTeamChain][4];  // world origin
    T tg[36];         // target poses of the three tasks: R (9) + p (3) each
    T J[12][16];      // weighted task Jacobian rows: slots 0-5 free-flyer columns, 6-12 the chain joints of the row's limb
    T e[16];          // weighted task error
    T piv[2][16];     // LDL^T: column k of the trailing matrix (double-buffered)
    T L[13][12];      // LDL^T multipliers by row; row 12 = D^-1 L^-1 e
    T dqff[8];        // free-flyer part of the step
};

// Registers a lane carries from one iteration to the next.
template <typename T> struct TeamLane {
    T qff[7];      // base position + quaternion (x, y, z, w), replicated in every lane
    T qr;          // this lane's revolute coordinate (joint 2 + lane, idx_q 7 + lane)
    T lo, hi;      // its limits
    T wgt;         // weight of task row `lane` (lanes 0-11)
};

template <typename T> IKB_HD T sel3(int i, T a, T b, T c) { return i == 0 ? a : (i == 1 ? b : c); }

// One ik::dls iteration (dls.cpp:16-71) of one problem by its 16 lanes.  Returns ||e||^2 of the priority-0 rows
// (visitor.hpp:19; identical in all lanes).  If that is below `tol` the state is left untouched (dls.cpp:61-64), else the
// lane's share of q has been stepped and clamped.
template <typename T, class SYNC>
IKB_HD T team_iteration(const int lane, TeamLane<T> &st, TeamScratch<T> &S, const TeamConsts<T> &C, const T step, const T damping2,
                        const T tol, SYNC &sync) {
    const int task = lane < 6 ? 0 : (lane < 9 ? 1 : (lane < 12 ? 2 : 0));
    const int comp = lane % 3;
    const int limb = task == 2 ? 1 : 0;
    const bool foot = lane >= 6 && lane < 12;
    const bool ang = lane >= 3 && lane < 6;

    // ---- phase 0: joint sin/cos, base rotation ----
    T Rb[9];
    quat_to_rot(st.qff[3], st.qff[4], st.qff[5], st.qff[6], Rb);
    {
        T s, c;
        sincos_(st.qr, &s, &c);
        S.sc[lane][0] = s;
        S.sc[lane][1] = c;
    }
    sync();

    // ---- phase 1: forward kinematics of my limb, row `comp` ----
    {
        T r0 = sel3(comp, Rb[0], Rb[3], Rb[6]), r1 = sel3(comp, Rb[1], Rb[4], Rb[7]), r2 = sel3(comp, Rb[2], Rb[5], Rb[8]);
        T pp = sel3(comp, st.qff[0], st.qff[1], st.qff[2]);
        const T b0 = r0, b1 = r1, b2 = r2, bp = pp;
#pragma unroll
        for (int k = 0; k < kTeamChain; ++k) {
            const T *PR = C.PR[limb][k], *Pp = C.Pp[limb][k];
            const T a0 = r0 * PR[0] + r1 * PR[3] + r2 * PR[6];
            const T a1 = r0 * PR[1] + r1 * PR[4] + r2 * PR[7];
            const T a2 = r0 * PR[2] + r1 * PR[5] + r2 * PR[8];
            pp = pp + (r0 * Pp[0] + r1 * Pp[1] + r2 * Pp[2]);
            const int jl = limb * 8 + (k < 6 ? k : 7);  // lane that owns chain joint k (slot 6 is the off-chain spring joint)
            const T s = S.sc[jl][0], c = S.sc[jl][1];
            r0 = c * a0 + s * a1;                        // A * Rz(q): columns 0, 1 mix, column 2 (the axis) is unchanged
            r1 = c * a1 - s * a0;
            r2 = a2;
            if (foot) {
                S.z[limb][k][comp] = a2;
                S.p[limb][k][comp] = pp;
            }
        }
        const T *FR = C.FR[limb], *Fp = C.Fp[limb];
        const T f0 = r0 * FR[0] + r1 * FR[3] + r2 * FR[6];
        const T f1 = r0 * FR[1] + r1 * FR[4] + r2 * FR[7];
        const T f2 = r0 * FR[2] + r1 * FR[5] + r2 * FR[8];
        const T fp = pp + (r0 * Fp[0] + r1 * Fp[1] + r2 * Fp[2]);
        if (foot || lane < 3) {  // the base task frame is the free-flyer joint frame itself
            T *F = S.frame[task];
            F[3 * comp + 0] = foot ? f0 : b0;
            F[3 * comp + 1] = foot ? f1 : b1;
            F[3 * comp + 2] = foot ? f2 : b2;
            F[9 + comp] = foot ? fp : bp;
        }
    }
    sync();

    // ---- phase 2: task error and Jlog6 blocks of my task; my rows of M1 = A Rf^T, M2 = B Rf^T ----
    T mA[3], mB[3], pf[3];
    {
        T Rf[9], Rt[9], pt[3];
        const T *F = S.frame[task], *G = S.tg + 12 * task;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            Rf[k] = F[k];
            Rt[k] = G[k];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            pf[k] = F[9 + k];
            pt[k] = G[9 + k];
        }
        // fMt = oMf^-1 * oMt (frame.hpp:48-50; `universe` reference: oMt = target)
        T Re[9], d[3], pe[3], w[3], th, sth, cth, lin[3];
        mat3T_mul(Rf, Rt, Re);
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = pt[k] - pf[k];
        rotT_vec(Rf, d, pe);
        log3(Re, w, th, sth, cth);
        const LogCoeffs<T> lc = log_coeffs(th, sth, cth);
        log6_from(w, lc, pe, lin);
        // tMf = fMt^-1: rotation Re^T (log3 = -w, same angle), translation Rt^T (pf - pt)   (frame.hpp:160-166)
        T nw[3] = {-w[0], -w[1], -w[2]}, nd[3] = {-d[0], -d[1], -d[2]}, p2[3], A[9], B[9];
        rotT_vec(Rt, nd, p2);
        jlog6_blocks(nw, th, lc, p2, A, B);
        const T ei = st.wgt * (ang ? sel3(comp, w[0], w[1], w[2]) : sel3(comp, lin[0], lin[1], lin[2]));  // data.cpp:49
        if (lane < 12) S.e[lane] = ei;
        T ar[3], br[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            ar[j] = sel3(comp, A[j], A[3 + j], A[6 + j]);
            br[j] = sel3(comp, B[j], B[3 + j], B[6 + j]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            mA[k] = ar[0] * Rf[3 * k] + ar[1] * Rf[3 * k + 1] + ar[2] * Rf[3 * k + 2];
            mB[k] = br[0] * Rf[3 * k] + br[1] * Rf[3 * k + 1] + br[2] * Rf[3 * k + 2];
        }
    }

    // ---- phase 3: my row of the weighted task Jacobian ----
    T jrow[13];
    {
        const T dpf[3] = {st.qff[0] - pf[0], st.qff[1] - pf[1], st.qff[2] - pf[2]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {  // free-flyer columns: world columns [R e_k; 0] and [p x R e_k; R e_k]
            const T rk[3] = {Rb[k], Rb[3 + k], Rb[6 + k]};
            T cx[3];
            cross3(dpf, rk, cx);
            const T t1 = dot3(mA, rk);
            const T t2 = dot3(mA, cx) + dot3(mB, rk);
            jrow[k] = ang ? T(0) : -t1;
            jrow[3 + k] = ang ? -t1 : -t2;
        }
#pragma unroll
        for (int j = 0; j < kTeamChain; ++j) {  // revolute columns of my limb: [(p_j - p_f) x z_j; z_j]
            const T *zj = S.z[limb][j], *pj = S.p[limb][j];
            const T z[3] = {zj[0], zj[1], zj[2]};
            const T dp[3] = {pj[0] - pf[0], pj[1] - pf[1], pj[2] - pf[2]};
            T cx[3];
            cross3(dp, z, cx);
            const T v = -(dot3(mA, cx) + dot3(mB, z));
            jrow[6 + j] = foot ? v : T(0);
        }
#pragma unroll
        for (int k = 0; k < 13; ++k) jrow[k] *= st.wgt;  // data.cpp:50
        if (lane < 12) {
#pragma unroll
            for (int k = 0; k < 13; ++k) S.J[lane][k] = jrow[k];
        }
    }
    sync();

    // ---- phase 4: my row of J J^T + damping^2 I; lane 12 carries the right-hand side ----
    T g[12];
    T res = T(0);
    {
#pragma unroll
        for (int b = 0; b < 12; ++b) {
            const T *Jb = S.J[b];
            T acc = jrow[0] * Jb[0];
#pragma unroll
            for (int k = 1; k < 6; ++k) acc += jrow[k] * Jb[k];
            g[b] = acc;
        }
        T cp[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {  // the chain columns are shared only with the rows of my own task
            const T *Jb = S.J[6 + 3 * limb + i];
            T acc = jrow[6] * Jb[6];
#pragma unroll
            for (int j = 1; j < kTeamChain; ++j) acc += jrow[6 + j] * Jb[6 + j];
            cp[i] = acc;
        }
#pragma unroll
        for (int b = 6; b < 12; ++b) g[b] += (foot && ((b >= 9) == (limb == 1))) ? cp[b % 3] : T(0);
#pragma unroll
        for (int b = 0; b < 12; ++b) g[b] += (b == lane) ? damping2 : T(0);
#pragma unroll
        for (int b = 0; b < 12; ++b) {
            const T eb = S.e[b];
            res += eb * eb;  // visitor.hpp:19
            if (lane == 12) g[b] = eb;
        }
    }

    // ---- phase 5: LDL^T by rows (no pivoting: the matrix is SPD thanks to the damping) ----
    T lm[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        T *pv = S.piv[k & 1];
        pv[lane] = g[k];
        sync();
        const T inv = rcp_(pv[k]);
        const T f = g[k] * inv;
        lm[k] = f;
#pragma unroll
        for (int b = k + 1; b < 12; ++b) g[b] -= f * pv[b];
    }

    // ---- phase 6: y = L^-T (D^-1 L^-1 e), replicated ----
    T y[12];
    if (lane < 13) {
#pragma unroll
        for (int k = 0; k < 12; ++k) S.L[lane][k] = lm[k];
    }
    sync();
#pragma unroll
    for (int k = 0; k < 12; ++k) y[k] = S.L[12][k];
#pragma unroll
    for (int a = 11; a >= 1; --a) {
#pragma unroll
        for (int k = 0; k < a; ++k) y[k] -= S.L[a][k] * y[a];
    }

    // ---- phase 7: dq = -J^T y ----
    T dqr;
    {
        const int jl = lane & 7, lb = lane >> 3;
        const int slot = 6 + (jl < 6 ? jl : 6);
        const T y0 = lb ? y[9] : y[6], y1 = lb ? y[10] : y[7], y2 = lb ? y[11] : y[8];
        const int r0 = 6 + 3 * lb;
        const T v = S.J[r0][slot] * y0 + S.J[r0 + 1][slot] * y1 + S.J[r0 + 2][slot] * y2;
        dqr = jl == 6 ? T(0) : -v;  // no task frame hangs below the spring joint
        if (lane < 6) {
            T acc = S.J[0][lane] * y[0];
#pragma unroll
            for (int a = 1; a < 12; ++a) acc += S.J[a][lane] * y[a];
            S.dqff[lane] = -acc;
        }
    }
    sync();

    // ---- phase 8: integrate and clamp (dls.cpp:61-71) ----
    if (!(res < tol)) {
        T v6[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) v6[k] = step * S.dqff[k];
        integrate_freeflyer(Rb, &st.qff[0], &st.qff[3], v6);
#pragma unroll
        for (int k = 0; k < 7; ++k) st.qff[k] = min_(C.upper[k], max_(st.qff[k], C.lower[k]));
        st.qr = min_(st.hi, max_(st.qr + step * dqr, st.lo));
    }
    return res;
}

#if defined(__CUDACC__)
// ---- device side: persistent teams, tickets as in dls_spec.cuh ------------------------------------------------
template <typename T> struct TeamLaunch {
    static constexpr int kTeamsPerCta = 8;  // 128 threads
    static constexpr int kCtasPerSm = 3;    // 24 teams per SM, <= 168 registers per thread
    static constexpr size_t kConstBytes = (sizeof(TeamConsts<T>) + 15) / 16 * 16;
    static constexpr size_t kSmem = kConstBytes + kTeamsPerCta * sizeof(TeamScratch<T>);
};

template <typename T, bool SEG>
__global__ void __launch_bounds__(TeamLaunch<T>::kTeamsPerCta *kTeamLanes, TeamLaunch<T>::kCtasPerSm)
    dls_team_kernel(const __grid_constant__ TeamConsts<T> gc, const __grid_constant__ SolveArgs<T> a) {
    extern __shared__ __align__(16) unsigned char team_smem[];
    TeamConsts<T> &C = *reinterpret_cast<TeamConsts<T> *>(team_smem);
    {
        const T *src = reinterpret_cast<const T *>(&gc);
        T *dst = reinterpret_cast<T *>(team_smem);
        for (int i = threadIdx.x; i < (int)(sizeof(TeamConsts<T>) / sizeof(T)); i += blockDim.x) dst[i] = src[i];
    }
    const int lane = threadIdx.x & (kTeamLanes - 1), team = threadIdx.x / kTeamLanes;
    TeamScratch<T> &S = reinterpret_cast<TeamScratch<T> *>(team_smem + TeamLaunch<T>::kConstBytes)[team];
    // a team without a problem keeps iterating on this harmless state (its lanes must take part in every __syncwarp)
    for (int i = lane; i < 36; i += kTeamLanes) S.tg[i] = (i % 12 == 0 || i % 12 == 4 || i % 12 == 8) ? T(1) : T(0);
    __syncthreads();

    TeamLane<T> st;
#pragma unroll
    for (int k = 0; k < 7; ++k) st.qff[k] = k == 6 ? T(1) : T(0);
    st.qr = T(0);
    st.lo = C.lower[7 + lane];
    st.hi = C.upper[7 + lane];
    st.wgt = C.weight[lane < 12 ? lane : 0];

    auto sync = []() { __syncwarp(); };
    long long b = 0;
    int it = 0;
    bool have = false, need = true;
    for (;;) {
        // every team that needs a problem pulls a ticket (lane 0) and loads it; warp-uniform control flow
        unsigned long long t = 0;
        if (need && lane == 0) t = atomicAdd(a.ticket, 1ULL);
        t = __shfl_sync(0xffffffffu, t, 0, kTeamLanes);
        if (need) {
            it = 0;
            if (!a.resume) {
                have = (long long)t < a.B;
                b = (long long)t;
            } else {
                have = t < *a.list_count;
                if (have) {
                    b = a.list[t];
                    it = a.iters_ws[b];
                }
            }
            if (have) {
                const ProblemIO<T> io = problem_io<SEG>(a, b);
                const T *qb = a.resume ? io.q : io.q0;
                const long long es = a.resume ? io.q_es : io.q0_es;
#pragma unroll
                for (int k = 0; k < 7; ++k) st.qff[k] = qb[k * es];
                st.qr = qb[(7 + lane) * es];
                for (int i = lane; i < 36; i += kTeamLanes) S.tg[i] = io.targets[i * io.tg_es];
            }
            need = false;
        }
        if (!__any_sync(0xffffffffu, have)) break;

        const T res = team_iteration(lane, st, S, C, a.step_length, a.damping2, a.tolerance, sync);

        if (have) {
            const bool converged = res < a.tolerance;      // visitor.hpp:19
            if (!converged) ++it;                           // team_iteration has stepped q
            if (converged || it >= a.max_iterations) {      // dls.cpp:61-64 / 14,76-77
                const ProblemIO<T> io = problem_io<SEG>(a, b);
#pragma unroll
                for (int k = 0; k < 7; ++k)
                    if (lane == k) io.q[k * io.q_es] = st.qff[k];
                io.q[(7 + lane) * io.q_es] = st.qr;
                if (lane == 0) {
                    if (io.success) *io.success = converged ? 1 : 0;
                    if (io.iters) *io.iters = it;
                    if (io.resid) *io.resid = res;
                }
                need = true;
                have = false;
            }
        }
    }
}

template <typename T, bool SEG> int launch_team_seg(const TeamConsts<T> &c, const SolveArgs<T> &a, long long n, int sm_count, cudaStream_t s) {
    using L = TeamLaunch<T>;
    auto fn = dls_team_kernel<T, SEG>;
    static DynSmemOptIn opt_in;  // per instantiation; the attribute itself is per device (dls_spec.cuh)
    if (!opt_in.ensure(fn, (int)L::kSmem)) return 1;
    long long ctas = (n + L::kTeamsPerCta - 1) / L::kTeamsPerCta;
    if (ctas > (long long)L::kCtasPerSm * sm_count) ctas = (long long)L::kCtasPerSm * sm_count;
    if (ctas < 1) ctas = 1;
    fn<<<(unsigned)ctas, L::kTeamsPerCta * kTeamLanes, L::kSmem, s>>>(c, a);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
template <typename T> int launch_team(const TeamConsts<T> &c, const SolveArgs<T> &a, long long n, int sm_count, cudaStream_t s) {
    return a.nseg > 0 ? launch_team_seg<T, true>(c, a, n, sm_count, s) : launch_team_seg<T, false>(c, a, n, sm_count, s);
}
#endif  // __CUDACC__

}  // namespace ikb
