// TEST-ONLY: compiles the GENERATED solver bodies (ik_b200/csrc/gen/*.cuh -- the very source the CUDA kernel in
// dls_spec.cuh instantiates) with g++ and drives them with the same loop as the kernel, so the generated arithmetic
// can be checked against the oracle on the GPU-less build box.  Never linked into libikb200.so; the product has no
// CPU path.
#include <barrier>
#include <thread>
#include <vector>

#include "../../ik_b200/csrc/gen/cassie_feet_pelvis.cuh"
#include "../../ik_b200/csrc/gen/cassie_feet_pelvis_arrow.cuh"
#include "../../ik_b200/csrc/gen/cassie_feet_pelvis_arrow_b.cuh"
#include "../../ik_b200/csrc/gen/cassie_feet_pelvis_w1.cuh"
#include "../../ik_b200/csrc/gen/cassie_feet_pelvis_w2.cuh"
#include "../../ik_b200/csrc/gen/humanoid_limbs.cuh"
#include "../../ik_b200/csrc/gen/humanoid_limbs_arrow.cuh"
#include "../../ik_b200/csrc/gen/manipulator_tool.cuh"
#include "../../ik_b200/csrc/gen/cassie_demo.cuh"
#include "../../ik_b200/csrc/gen/cassie_demo_posture.cuh"

using namespace ikb;

template <class Spec, typename T>
static int spec_solve(const double *lower, const double *upper, const double *weight, const double *mask, const double *q0,
                      const double *targets, int max_it, double step, double damping, double tol, double *q_out, int *iters,
                      double *resid, double *e_first, bool parallel = false) {
    constexpr int NQ = Spec::NQ, NV = Spec::NV, M = Spec::M;
    if (Spec::ARROW) parallel = true;  // the arrow specs have no serial solve (their strip has no room for a dense factor)
    SpecConsts<T, NQ, M> c;
    for (int k = 0; k < NQ; ++k) { c.lower[k] = (T)lower[k]; c.upper[k] = (T)upper[k]; }
    for (int i = 0; i < M; ++i) {
        c.weight[i] = (T)weight[i];
        c.mask[i] = mask ? (T)mask[i] : T(1);
    }
    std::vector<T> bufJ(Spec::NSLOT), bufL(Spec::NFACT), bufT(Spec::TSZ), tgt(targets, targets + Spec::TSZ);
    const Strip<T, 1> sJ{bufJ.data()}, sL{bufL.data()}, sT{bufT.data()};
    const Strip<T, 1> sE{bufL.data() + Spec::EOFF};  // e aliases part of the factor strip, as in the kernel
    for (int role = 0; role < Spec::NWARPS; ++role) Spec::load_targets(role, tgt.data(), 1LL, sT);
    T q[NQ];
    for (int k = 0; k < NQ; ++k) q[k] = (T)q0[k];
    // distributed step (Spec::DSTEP): every role keeps its own copy of q and steps only the coordinates it owns, as in the kernel
    const bool dstep = parallel && Spec::NWARPS > 1 && Spec::DSTEP;
    constexpr int NQL = Spec::NQL;   // (arrow specs: a role keeps only the coordinates it reads and steps)
    std::vector<T> qr_buf((size_t)Spec::NWARPS * NQL);
    auto qr = [&](int role) -> T(&)[NQL] { return *reinterpret_cast<T(*)[NQL]>(qr_buf.data() + (size_t)role * NQL); };
    for (int role = 0; role < Spec::NWARPS; ++role) Spec::load_q(role, q, 1LL, qr(role));
    int it = 0, success = 0;
    T res = 0;
    while (it < max_it) {
        for (int role = 0; role < Spec::NWARPS; ++role) {   // the warp roles, in turn
            if constexpr (NQL == NQ) Spec::evaluate(role, dstep ? qr(role) : q, sT, c, sJ, sE);
            else Spec::evaluate(role, qr(role), sT, c, sJ, sE);
        }
        if (Spec::PRE > 0) Spec::presolve(sJ, sL, sE, (T)(damping * damping));                      // solver role, before the barrier
        if (it == 0 && e_first) for (int i = 0; i < M; ++i) e_first[i] = (double)sE.get(i);
        T y[M], dq[NV];
        std::vector<T> yr_buf((size_t)Spec::NWARPS * Spec::MY);   // arrow specs with y in the roles' registers
        auto yr = [&](int role) -> T(&)[Spec::MY] { return *reinterpret_cast<T(*)[Spec::MY]>(yr_buf.data() + (size_t)role * Spec::MY); };
        if (parallel && Spec::NWARPS > 1) {
            // the role-distributed solve: one host thread per warp role, a std::barrier as the group barrier
            res = 0;
            for (int i = 0; i < Spec::M0; ++i) res += sE.get(i) * sE.get(i);
            std::barrier<> bar(Spec::NWARPS);
            std::vector<std::thread> th;
            for (int role = 0; role < Spec::NWARPS; ++role)
                th.emplace_back([&, role]() {
                    T yl[M];
                    auto sync = [&]() { bar.arrive_and_wait(); };
                    if constexpr (Spec::ARROW) {
                        auto hook = []() {};
                        Spec::psolve(role, sJ, sL, sE, (T)(damping * damping), yr(role), sync, hook);
                    } else {
                        Spec::psolve(role, sJ, sL, sE, (T)(damping * damping), yl, sync);
                    }
                    if (role == Spec::SOLVER) for (int i = 0; i < M; ++i) y[i] = yl[i];
                });
            for (auto &t : th) t.join();
        } else {
            if constexpr (!Spec::ARROW) res = Spec::solve(sJ, sL, sE, (T)(damping * damping), y);
        }
        if (res < (T)tol) { success = 1; break; }
        if constexpr (Spec::DSTEP) {
            if (dstep) {
                for (int role = 0; role < Spec::NWARPS; ++role) {
                    if constexpr (Spec::MY != M) Spec::step_role(role, sJ, sL, qr(role), (T)step, c, yr(role));
                    else Spec::step_role(role, sJ, sL, qr(role), (T)step, c);
                }
                if constexpr (Spec::QCOMMON)   // (the kernel does this behind the barrier at the top of the next trip)
                    for (int role = 0; role < Spec::NWARPS; ++role) Spec::fetch_common(role, sL, qr(role));
                ++it;
                continue;
            }
        }
        if constexpr (NQL == NQ) {
            Spec::step_direction(sJ, y, dq);
            Spec::integrate(q, dq, (T)step, c);
        }
        ++it;
    }
    if constexpr (Spec::DSTEP) {
        if (dstep)
            for (int role = 0; role < Spec::NWARPS; ++role) Spec::store_q(role, qr(role), q, 1LL);  // the result, assembled as the kernel does
    }
    for (int k = 0; k < NQ; ++k) q_out[k] = (double)q[k];
    *iters = it;
    *resid = (double)res;
    return success;
}

// one evaluation: weighted error e[M] and the dense weighted task Jacobian J[M][NV] rebuilt from the strip
template <class Spec>
static void spec_eval(const double *weight, const double *q0, const double *targets, double *e_out, double *J_out) {
    constexpr int NQ = Spec::NQ, NV = Spec::NV, M = Spec::M;
    SpecConsts<double, NQ, M> c{};
    for (int i = 0; i < M; ++i) {
        c.weight[i] = weight[i];
        c.mask[i] = 1.0;
    }
    std::vector<double> bufJ(Spec::NSLOT), bufT(targets, targets + Spec::TSZ), bufE(M);
    const Strip<double, 1> sJ{bufJ.data()}, sT{bufT.data()}, sE{bufE.data()};
    for (int role = 0; role < Spec::NWARPS; ++role) {
        double q[Spec::NQL];
        Spec::load_q(role, q0, 1LL, q);
        Spec::evaluate(role, q, sT, c, sJ, sE);
    }
    for (int i = 0; i < M; ++i) e_out[i] = bufE[i];
    for (int i = 0; i < M * NV; ++i) J_out[i] = 0;
    for (int k = 0; k < Spec::NSLOT; ++k) J_out[Spec::slot_rc()[2 * k] * NV + Spec::slot_rc()[2 * k + 1]] = bufJ[k];
}

#define IKB_SPEC_EXPORT(fn, Spec)                                                                                        \
    extern "C" int fn##_d(const double *lo, const double *hi, const double *w, const double *mk, const double *q0,       \
                          const double *tg, int mi, double st, double da, double tol, double *q, int *it, double *res,  \
                          double *e0) {                                                                                  \
        return spec_solve<Spec, double>(lo, hi, w, mk, q0, tg, mi, st, da, tol, q, it, res, e0);                         \
    }                                                                                                                    \
    extern "C" int fn##_f(const double *lo, const double *hi, const double *w, const double *mk, const double *q0,       \
                          const double *tg, int mi, double st, double da, double tol, double *q, int *it, double *res,  \
                          double *e0) {                                                                                  \
        return spec_solve<Spec, float>(lo, hi, w, mk, q0, tg, mi, st, da, tol, q, it, res, e0);                          \
    }                                                                                                                    \
    extern "C" int fn##_pd(const double *lo, const double *hi, const double *w, const double *mk, const double *q0,      \
                           const double *tg, int mi, double st, double da, double tol, double *q, int *it, double *res, \
                           double *e0) {                                                                                 \
        return spec_solve<Spec, double>(lo, hi, w, mk, q0, tg, mi, st, da, tol, q, it, res, e0, true);                   \
    }                                                                                                                    \
    extern "C" void fn##_eval(const double *w, const double *q0, const double *tg, double *e, double *J) {               \
        spec_eval<Spec>(w, q0, tg, e, J);                                                                                \
    }

IKB_SPEC_EXPORT(h_spec_cassie_feet_pelvis, SpecCassieFeetPelvis)
IKB_SPEC_EXPORT(h_spec_cassie_feet_pelvis_arrow, SpecCassieFeetPelvisArrow)
IKB_SPEC_EXPORT(h_spec_cassie_feet_pelvis_arrow_b, SpecCassieFeetPelvisArrowB)
IKB_SPEC_EXPORT(h_spec_cassie_feet_pelvis_w1, SpecCassieFeetPelvisW1)
IKB_SPEC_EXPORT(h_spec_cassie_feet_pelvis_w2, SpecCassieFeetPelvisW2)
IKB_SPEC_EXPORT(h_spec_manipulator_tool, SpecManipulatorTool)
IKB_SPEC_EXPORT(h_spec_humanoid_limbs, SpecHumanoidLimbs)
IKB_SPEC_EXPORT(h_spec_humanoid_limbs_arrow, SpecHumanoidLimbsArrow)
IKB_SPEC_EXPORT(h_spec_cassie_demo, SpecCassieDemo)
IKB_SPEC_EXPORT(h_spec_cassie_demo_posture, SpecCassieDemoPosture)
