// The reference's only caller of ik::dls -- CassieIK::init / CassieIK::loop, ik_ros/src/cassie.cpp:19-130 -- without
// ROS, written against include/ik/*.hpp (same names, same calls).  Each tick moves the left-foot target
// (cassie.cpp:95-96), warm-starts from the previous solution (cassie.cpp:112) and solves with the demo parameters
// (damping 1e-1, 200 iterations, step 1e-1; cassie.cpp:105-109).
//
//   g++ -std=c++17 -Iinclude examples/cassie_ik_demo.cpp -Lik_b200 -likb200 -Wl,-rpath,$PWD/ik_b200 -o build/cassie_ik_demo
//   build/cassie_ik_demo ik_b200/data/cassie.urdf [ticks]
//
// Prints one line per tick: tick, success, iterations, ||e||^2, q[7..10]; and a final "batch" line from ik::dls_batch.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>

#include "ik/dls.hpp"
#include "ik/frame.hpp"
#include "ik/centre_of_mass.hpp"
#include "ik/pik.hpp"
#include "ik/problem.hpp"

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <cassie.urdf> [ticks]\n", argv[0]);
        return 2;
    }
    const int ticks = argc > 2 ? std::atoi(argv[2]) : 10;
    std::ifstream f(argv[1]);
    std::stringstream ss;
    ss << f.rdbuf();
    try {
        ik::model_t model;
        ik::urdf::buildModelFromXML(ss.str(), /*free_flyer=*/true, model);  // cassie.cpp:34-35
        ik::InverseKinematicsProblem problem(model, 1);                      // cassie.cpp:43

        auto fl = ik::FrameTask::create(model, "LeftFootFront", ik::KinematicType::Position, "pelvis");  // cassie.cpp:45-46
        auto pelvis = ik::FrameTask::create(model, "pelvis", ik::KinematicType::Full);                   // cassie.cpp:54
        auto align = ik::AlignAxisTask::create(model, "LeftFootFront", ik::AlignAxisType::AxisY);        // cassie.cpp:60-62
        align->target = {1.0, 0.0, 0.0};
        problem.add_frame_task("fl", fl);          // cassie.cpp:73
        problem.add_frame_task("pelvis", pelvis);  // cassie.cpp:77
        problem.add_align_axis_task("fl_align", align);  // cassie.cpp:81

        ik::vector_t q = model.neutral();  // zeros, q[6] = 1 (cassie.cpp:68-70)
        ik::dls_data data(problem);        // cassie.cpp:86
        ik::dls_parameters p;
        p.damping = 1e-1;
        p.max_iterations = 200;
        p.step_length = 1e-1;
        for (int k = 0; k < ticks; ++k) {
            const double t = 0.02 * k;
            problem.get_frame_task("fl")->target.translation() = {0.0, 0.1, -0.6 + 0.2 * std::sin(0.5 * t)};  // cassie.cpp:95-96
            problem.get_frame_task("pelvis")->target = ik::se3_t::Identity();                                // cassie.cpp:98-99
            q = ik::dls(problem, q, data, ik::inverse_kinematics_visitor(), p);                              // cassie.cpp:112
            std::printf("tick %d success %d iterations %d resid %.12e q7..10 %.12e %.12e %.12e %.12e\n", k, (int)data.success,
                        data.info.iterations, data.residual, q[7], q[8], q[9], q[10]);
        }
        // batched extension: the same problem for 4 foot heights at once
        const int B = 4, nq = model.nq;
        std::vector<double> q0(B * nq), tg;
        for (int b = 0; b < B; ++b) {
            std::copy(q.begin(), q.end(), q0.begin() + b * nq);
            problem.get_frame_task("fl")->target.translation() = {0.0, 0.1, -0.7 + 0.05 * b};
            const ik::vector_t t = problem.gather_targets();
            tg.insert(tg.end(), t.begin(), t.end());
        }
        const ik::dls_batch_result r = ik::dls_batch(problem, B, q0.data(), tg.data(), ik::inverse_kinematics_visitor(), p);
        for (int b = 0; b < B; ++b)
            std::printf("batch %d success %d iterations %d resid %.12e\n", b, (int)r.success[b], r.iterations[b], r.residual[b]);
        // stream of batches: three copies of that batch through the pipelined queue, merged into one kernel pair
        ik::dls_batch_queue queue(problem, /*depth=*/4, /*merge=*/3);
        ik::dls_batch_result rq[3];
        ik::dls_batch_queue::ticket_t tk[3];
        for (int k = 0; k < 3; ++k) tk[k] = queue.submit(B, q0.data(), tg.data(), rq[k], ik::inverse_kinematics_visitor(), p);
        int same = 0;
        for (int k = 0; k < 3; ++k) {
            queue.wait(tk[k]);
            same += rq[k].q == r.q && rq[k].iterations == r.iterations && rq[k].success == r.success;
        }
        std::printf("queue: %d of 3 merged batches identical to ik::dls_batch\n", same);
        // several GPUs behind the same call (ikb_multi_*): the batch is cut into one slice per listed device.  With one GPU
        // in the box the two slices share it -- the sharding logic is the same.
        const int ndev = ikb_device_count();
        const ik::dls_batch_result rm = ik::dls_batch(problem, B, q0.data(), tg.data(), std::vector<int>{0, ndev > 1 ? 1 : 0},
                                                      ik::inverse_kinematics_visitor(), p);
        std::printf("multi: devices %d sharded batch identical %d\n", ndev, (int)(rm.q == r.q && rm.iterations == r.iterations && rm.success == r.success));
        // dls_data after a solve (data.hpp:15-28): dq, e, J of the last evaluation; and a visitor that overrides should_stop
        {
            struct step_visitor : ik::inverse_kinematics_visitor {   // stop when the step is small instead of the error
                bool should_stop(const ik::InverseKinematicsProblem &, const std::vector<ik::vector_t> &, const ik::vector_t &dq) const override {
                    double m = 0;
                    for (double x : dq) m = std::max(m, std::fabs(x));
                    return m < 5e-2;
                }
                bool is_default_test() const override { return false; }
            };
            ik::dls_data d1(problem), d2(problem);
            problem.get_frame_task("fl")->target.translation() = {0.0, 0.1, -0.6};
            ik::dls(problem, model.neutral(), d1, ik::inverse_kinematics_visitor(), p);
            ik::dls(problem, model.neutral(), d2, step_visitor(), p);
            double e2 = 0, dqmax = 0;
            for (double x : d1.e[0]) e2 += x * x;
            for (double x : d2.dq) dqmax = std::max(dqmax, std::fabs(x));
            std::printf("data: e_rows %d J_entries %d |e|^2-resid %.3e iterations %d | custom stop: success %d iterations %d max|dq| %.6e\n",
                        (int)d1.e[0].size(), (int)d1.J[0].size(), std::fabs(e2 - d1.residual), d1.info.iterations, (int)d2.success,
                        d2.info.iterations, dqmax);
        }
        // the demo's other solver (IKMethod::PIK, cassie.cpp:114-124): one tick with ik::pik and the demo's parameters
        ik::pik_data pdata(problem);  // lambda = 1.0 per priority level (pik.hpp:31)
        ik::pik_parameters pp;
        pp.damping = 1e-2;
        pp.max_iterations = 200;
        pp.step_length = 1e0;
        problem.get_frame_task("fl")->target.translation() = {0.0, 0.1, -0.6};
        const ik::vector_t qp = ik::pik(problem, model.neutral(), pdata, ik::inverse_kinematics_visitor(), pp);
        std::printf("pik success %d iterations %d resid %.12e q7..10 %.12e %.12e %.12e %.12e\n", (int)pdata.success,
                    pdata.info.iterations, pdata.residual, qp[7], qp[8], qp[9], qp[10]);
        // the demo's commented-out contact constraint (cassie.cpp:49-51,75: "keep the foot in place"): right foot pinned
        ik::InverseKinematicsProblem pinned(model, 0);
        auto pel = ik::FrameTask::create(model, "pelvis", ik::KinematicType::Full);
        pel->target.translation() = {0.0, 0.01, -0.02};
        pinned.add_frame_task("pelvis", pel);
        pinned.add_frame_constraint("fr", ik::FrameConstraint::create(model, "RightFootFront", ik::KinematicType::Full));
        ik::dls_data cdata(pinned);
        ik::dls_parameters cp;
        cp.step_length = 0.5;
        const ik::vector_t qc = ik::dls(pinned, model.neutral(), cdata, ik::inverse_kinematics_visitor(), cp);
        std::printf("constraint c_size %d success %d iterations %d resid %.12e q7..10 %.12e %.12e %.12e %.12e\n", (int)pinned.c_size(),
                    (int)cdata.success, cdata.info.iterations, cdata.residual, qc[7], qc[8], qc[9], qc[10]);
        // CentreOfMassTask (centre_of_mass.hpp:14-52): shift the centre of mass, keep the pelvis upright
        ik::InverseKinematicsProblem balance(model, 0);
        auto com = ik::CentreOfMassTask::create(model);
        com->target = {0.02, 0.01, -0.25};
        balance.add_centre_of_mass_task(com);
        balance.add_frame_task("pelvis", ik::FrameTask::create(model, "pelvis", ik::KinematicType::Orientation));
        ik::dls_data bdata(balance);
        const ik::vector_t qb = ik::dls(balance, model.neutral(), bdata, ik::inverse_kinematics_visitor(), cp);
        std::printf("com e_size %d success %d iterations %d resid %.12e q0..3 %.12e %.12e %.12e %.12e\n", (int)balance.e_size(0),
                    (int)bdata.success, bdata.info.iterations, bdata.residual, qb[0], qb[1], qb[2], qb[7]);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
