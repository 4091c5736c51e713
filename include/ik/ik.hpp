// ik/ik.hpp -- C++ facade with the reference's task / solver surface (namespace ik), over the C ABI of libikb200.so.
//
// Drop-in for the damped-least-squares path of dazzmo/ik: the same class and function names, argument meaning and
// failure behaviour (a failed solve = data.success == false + the last iterate, reference ik/ik/dls.cpp:76-77), with the
// arithmetic running in the CUDA kernels behind include/ikb200.h.  Header-only; link with -likb200.
//
// Reference interface                                   (file:line)                    here
//   ik::number_t / index_t / string_t                    ik/ik/common.hpp:11-15         same typedefs
//   ik::model_t  (= pinocchio::Model)                    ik/ik/common.hpp:17            ik::model_t (flat tree handle)
//   pinocchio::urdf::buildModelFromXML(xml, FF(), m)     ik_ros/src/cassie.cpp:34-35    ik::urdf::buildModelFromXML
//   ik::se3_t    (= pinocchio::SE3)                      ik/ik/common.hpp:20            ik::se3_t {rotation(), translation()}
//   ik::vector_t (= Eigen::VectorXd)                     ik/ik/common.hpp:28            ik::vector_t (std::vector<double>)
//   ik::default_solver_parameters                        ik/ik/common.hpp:59-66         same fields, same defaults
//   ik::Task::dimension() / weighting()                  ik/ik/task.hpp:19-57           same
//   ik::KinematicType, ik::FrameTask (+ ::create, target) ik/ik/frame.hpp:20,78-200     same
//   ik::AlignAxisType, ik::AlignAxisTask                 ik/ik/frame.hpp:202-319        same
//   ik::PostureTask                                      ik/ik/posture.hpp:17-86        same
//   ik::CentreOfMassTask                                 ik/ik/centre_of_mass.hpp:14-52 same (add_/get_centre_of_mass_task)
//   ik::InverseKinematicsProblem                         ik/ik/problem.hpp:9-206        same members (frame/axis/posture tasks)
//   ik::inverse_kinematics_visitor                       ik/ik/visitor.hpp:7-24         default stop test; `tolerance` member
//   ik::dls_parameters, ik::dls_data, ik::dls_info       ik/ik/dls.hpp:24-74            same (+ iterations / residual filled)
//   ik::vector_t ik::dls(problem, q0, data, visitor, p)  ik/ik/dls.hpp:111-114          same signature
//   -- extension the reference lacks --                                                 ik::dls_batch (host arrays), ik::dls_batch_queue;  ik::pik / pik_data / pik_parameters (pik.hpp)
//
//   -- several GPUs --                                                                  ik::dls_batch(..., devices): the batch is sharded over the listed GPUs (ikb_multi_*)
//
// Eigen is not a dependency: vector_t / se3_t are minimal value types with the accessors the reference's callers use.
// Where <Eigen/Core> is available (the reference's own environment, common.hpp:23-38) the header also provides the
// Eigen / Pinocchio-shaped overloads: dls(problem, Eigen::VectorXd, ...) -> Eigen::VectorXd, se3_t from / to
// (Eigen::Matrix3d, Eigen::Vector3d), dls_data::e_eigen() / J_eigen() / dq_eigen() -- so the reference's call sites
// (cassie.cpp:95-113) compile unchanged.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../ikb200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && !defined(IK_NO_EIGEN)
#include <Eigen/Core>
#define IK_HAVE_EIGEN 1
#endif
#endif

namespace ik {

typedef int int_t;
typedef std::size_t index_t;
typedef double number_t;
typedef std::string string_t;
typedef std::vector<number_t> vector_t;

inline void check(int status, const char *what) {
    if (status != IKB_OK) throw std::runtime_error(std::string(what) + ": " + ikb_last_error());
}

struct se3_t {  // pinocchio::SE3: rotation row-major, translation
    std::array<number_t, 9> R{{1, 0, 0, 0, 1, 0, 0, 0, 1}};
    std::array<number_t, 3> p{{0, 0, 0}};
    static se3_t Identity() { return se3_t(); }
    std::array<number_t, 9> &rotation() { return R; }
    const std::array<number_t, 9> &rotation() const { return R; }
    std::array<number_t, 3> &translation() { return p; }
    const std::array<number_t, 3> &translation() const { return p; }
#ifdef IK_HAVE_EIGEN
    // pinocchio::SE3(R, p) / .rotation() / .translation() for Eigen callers (cassie.cpp:95-99 writes target.translation())
    se3_t() = default;
    se3_t(const Eigen::Matrix3d &Rm, const Eigen::Vector3d &pv) {
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) R[3 * i + j] = Rm(i, j);
            p[i] = pv[i];
        }
    }
    Eigen::Map<Eigen::Matrix<number_t, 3, 3, Eigen::RowMajor>> rotation_eigen() { return Eigen::Map<Eigen::Matrix<number_t, 3, 3, Eigen::RowMajor>>(R.data()); }
    Eigen::Map<Eigen::Matrix<number_t, 3, 1>> translation_eigen() { return Eigen::Map<Eigen::Matrix<number_t, 3, 1>>(p.data()); }
#endif
};

// Flattened kinematic tree; copies share the immutable handle (the reference copies the Pinocchio model by value).
class model_t {
   public:
    model_t() = default;
    explicit model_t(ikb_model *h) : h_(h, ikb_model_free) { refresh(); }
    int nq = 0, nv = 0, njoints = 0, nframes = 0;
    vector_t lowerPositionLimit, upperPositionLimit;
    // model.getFrameId(name): nframes when absent (reference ik/ik/common.hpp:50)
    index_t getFrameId(const string_t &name) const { return (index_t)ikb_model_frame_id(h_.get(), name.c_str()); }
    bool existFrame(const string_t &name) const { return getFrameId(name) < (index_t)nframes; }
    vector_t neutral() const {
        vector_t q(nq);
        check(ikb_model_neutral(h_.get(), q.data()), "ikb_model_neutral");
        return q;
    }
    const ikb_model *handle() const { return h_.get(); }

   private:
    void refresh() {
        nq = ikb_model_nq(h_.get());
        nv = ikb_model_nv(h_.get());
        njoints = ikb_model_njoints(h_.get());
        nframes = ikb_model_nframes(h_.get());
        lowerPositionLimit.resize(nq);
        upperPositionLimit.resize(nq);
        check(ikb_model_get_limits(h_.get(), lowerPositionLimit.data(), upperPositionLimit.data()), "ikb_model_get_limits");
    }
    std::shared_ptr<ikb_model> h_;
};

namespace urdf {
// pinocchio::urdf::buildModelFromXML(xml, pinocchio::JointModelFreeFlyer(), model)  (ik_ros/src/cassie.cpp:34-35);
// free_flyer = false is buildModelFromXML(xml, model).
inline model_t &buildModelFromXML(const string_t &xml, bool free_flyer, model_t &model) {
    ikb_model *h = nullptr;
    check(ikb_model_from_urdf(xml.data(), xml.size(), free_flyer ? 1 : 0, &h), "ikb_model_from_urdf");
    model = model_t(h);
    return model;
}
}  // namespace urdf

struct default_solver_parameters {  // ik/ik/common.hpp:59-66
    index_t max_iterations = 100;
    number_t max_time = 1.0;  // declared by the reference, never read
    number_t step_length = 1.0;
};

class Task {  // ik/ik/task.hpp:19-57
   public:
    enum class Kind { Frame, AlignAxis, Posture, CentreOfMass };
    Task() : dimension_(0) {}
    explicit Task(const index_t &dimension) { set_dimension(dimension); }
    virtual ~Task() = default;
    virtual Kind kind() const = 0;
    index_t dimension() const { return dimension_; }
    vector_t &weighting() { return weighting_; }
    const vector_t &weighting() const { return weighting_; }

   protected:
    void set_dimension(const index_t &dimension) {
        dimension_ = dimension;
        weighting_.assign(dimension, 1.0);
    }

   private:
    index_t dimension_;
    vector_t weighting_;
};

enum class KinematicType { Position = 0, Orientation, Full };  // ik/ik/frame.hpp:20

class FrameTask : public Task {  // ik/ik/frame.hpp:78-200
   public:
    FrameTask(const model_t &model, const std::string &frame, const KinematicType &type = KinematicType::Full,
              const std::string &reference_frame = "universe")
        : Task(), target(se3_t::Identity()), type(type), frame(frame), reference_frame(reference_frame) {
        (void)model;
        set_dimension(type == KinematicType::Full ? 6 : 3);
    }
    static std::shared_ptr<FrameTask> create(const model_t &model, const std::string &frame,
                                             const KinematicType &type = KinematicType::Full,
                                             const std::string &reference_frame = "universe") {
        return std::make_shared<FrameTask>(model, frame, type, reference_frame);
    }
    Kind kind() const override { return Kind::Frame; }
    se3_t target;  // public and mutated between solves, like the reference (frame.hpp:189, cassie.cpp:95-99)
    KinematicType type;
    std::string frame, reference_frame;
};

enum class AlignAxisType { AxisX = 0, AxisY, AxisZ };  // ik/ik/frame.hpp:202

class AlignAxisTask : public Task {  // ik/ik/frame.hpp:210-319
   public:
    AlignAxisTask(const model_t &model, const std::string &frame, const AlignAxisType &axis,
                  const std::string &reference_frame = "universe")
        : Task(1), target{{1, 0, 0}}, axis(axis), frame(frame), reference_frame(reference_frame) {
        (void)model;
    }
    static std::shared_ptr<AlignAxisTask> create(const model_t &model, const std::string &frame, const AlignAxisType &axis,
                                                 const std::string &reference_frame = "universe") {
        return std::make_shared<AlignAxisTask>(model, frame, axis, reference_frame);
    }
    Kind kind() const override { return Kind::AlignAxis; }
    std::array<number_t, 3> target;
    AlignAxisType axis;
    std::string frame, reference_frame;
};

// ik/ik/frame.hpp:333-465: a hard constraint -- ik::dls projects its step into the null space of the stacked constraint
// Jacobian (dls.cpp:26-34,44-52), i.e. keeps `frame` at rest relative to `reference_frame`.
class FrameConstraint {
   public:
    FrameConstraint(const model_t &model, const std::string &frame, const KinematicType &type = KinematicType::Full,
                    const std::string &reference_frame = "universe")
        : type(type), frame(frame), reference_frame(reference_frame) {
        (void)model;
    }
    static std::shared_ptr<FrameConstraint> create(const model_t &model, const std::string &frame,
                                                   const KinematicType &type = KinematicType::Full,
                                                   const std::string &reference_frame = "universe") {
        return std::make_shared<FrameConstraint>(model, frame, type, reference_frame);
    }
    index_t dimension() const { return type == KinematicType::Full ? 6 : 3; }
    se3_t target = se3_t::Identity();  // declared by the reference, not read by ik::dls
    KinematicType type;
    std::string frame, reference_frame;
};

class PostureTask : public Task {  // ik/ik/posture.hpp:17-86
   public:
    PostureTask(const model_t &model, const index_t &nj) : Task(nj), target(nj, 0.0), mask(nj, 1.0), nj(nj) { (void)model; }
    static std::shared_ptr<PostureTask> create(const model_t &model, const index_t &nj) {
        return std::make_shared<PostureTask>(model, nj);
    }
    Kind kind() const override { return Kind::Posture; }
    vector_t target, mask;
    index_t nj;
};

class CentreOfMassTask : public Task {  // ik/ik/centre_of_mass.hpp:14-52
   public:
    CentreOfMassTask(const model_t &model, const std::string &reference_frame) : Task(3), reference_frame(reference_frame) {
        (void)model;
    }
    static std::shared_ptr<CentreOfMassTask> create(const model_t &model, const std::string &reference_frame = "universe") {
        return std::make_shared<CentreOfMassTask>(model, reference_frame);
    }
    Kind kind() const override { return Kind::CentreOfMass; }
    std::array<number_t, 3> target{{0, 0, 0}};  // centre of mass in the task's reference frame
    std::string reference_frame;         // (private in the reference; the bridge to the C ABI reads it)
};

class InverseKinematicsProblem {  // ik/ik/problem.hpp:9-206
   public:
    InverseKinematicsProblem(const model_t &model, const std::size_t &max_priority_level = 0)
        : model_(model), max_priority_level_(max_priority_level), tasks_(max_priority_level + 1) {}
    ~InverseKinematicsProblem() { release(); }
    InverseKinematicsProblem(const InverseKinematicsProblem &) = delete;
    InverseKinematicsProblem &operator=(const InverseKinematicsProblem &) = delete;

    const std::size_t &max_priority_level() const { return max_priority_level_; }
    std::size_t e_size(const std::size_t &priority) const {
        std::size_t sz = 0;
        for (const auto &task : get_all_tasks(priority)) sz += task->dimension();
        return sz;
    }
    std::size_t c_size() const {  // problem.hpp:47-53
        std::size_t sz = 0;
        for (const auto &c : constraints_) sz += c->dimension();
        return sz;
    }
    std::shared_ptr<FrameConstraint> add_frame_constraint(const std::string &name, const std::shared_ptr<FrameConstraint> &c) {
        constraints_map_.insert({name, constraints_.size()});  // problem.hpp:107-118
        constraints_.push_back(c);
        release();
        return constraints_.back();
    }
    std::shared_ptr<FrameConstraint> get_frame_constraint(const std::string &name) { return constraints_.at(constraints_map_.at(name)); }
    const std::vector<std::shared_ptr<FrameConstraint>> &get_all_constraints() const { return constraints_; }

    std::shared_ptr<FrameTask> add_frame_task(const std::string &name, const std::shared_ptr<FrameTask> &task,
                                              const std::size_t &priority = 0) {
        return add(frame_tasks_map_, frame_tasks_, name, task, priority);
    }
    std::shared_ptr<FrameTask> get_frame_task(const std::string &name) { return frame_tasks_.at(frame_tasks_map_.at(name)); }
    std::shared_ptr<AlignAxisTask> add_align_axis_task(const std::string &name, const std::shared_ptr<AlignAxisTask> &task,
                                                       const std::size_t &priority = 0) {
        return add(axis_tasks_map_, axis_tasks_, name, task, priority);
    }
    // (the reference looks this name up in the FRAME-task map, problem.hpp:109; here it is the axis-task map)
    std::shared_ptr<AlignAxisTask> get_align_axis_task(const std::string &name) { return axis_tasks_.at(axis_tasks_map_.at(name)); }
    std::shared_ptr<PostureTask> add_posture_task(const std::string &name, const std::shared_ptr<PostureTask> &task,
                                                  const std::size_t &priority = 0) {
        return add(posture_tasks_map_, posture_tasks_, name, task, priority);
    }
    std::shared_ptr<PostureTask> get_posture_task(const std::string &name) { return posture_tasks_.at(posture_tasks_map_.at(name)); }
    std::shared_ptr<CentreOfMassTask> add_centre_of_mass_task(const std::shared_ptr<CentreOfMassTask> &task,
                                                              const std::size_t &priority = 0) {  // problem.hpp:121-128
        if (priority > max_priority_level_) throw std::out_of_range("Maximum priority level exceeded!");
        com_task_ = task;
        tasks_[priority].push_back(task);
        ordered_.emplace_back(task, priority);
        release();
        return com_task_;
    }
    std::shared_ptr<CentreOfMassTask> get_centre_of_mass_task() { return com_task_; }
    const std::vector<std::shared_ptr<Task>> &get_all_tasks(const std::size_t &priority) const { return tasks_.at(priority); }
    const model_t &model() const { return model_; }

    // ---- bridge to the C ABI (not part of the reference surface) ----
    // Tasks in insertion order, which is the order the ABI lays target data out in.
    const std::vector<std::pair<std::shared_ptr<Task>, std::size_t>> &ordered_tasks() const { return ordered_; }
    // Finalized C handle for `device`; rebuilt when tasks or weights changed since the last call.
    ikb_problem *handle(int device = 0) {
        vector_t w;   // everything finalize bakes into the handle: row weights and posture masks (data.cpp:49-50, posture.hpp:52)
        for (const auto &tp : ordered_) {
            for (number_t x : tp.first->weighting()) w.push_back(x);
            if (tp.first->kind() == Task::Kind::Posture)
                for (number_t x : static_cast<const PostureTask &>(*tp.first).mask) w.push_back(x);
        }
        if (h_ && w == baked_weights_ && device == device_) return h_;
        release();
        check(ikb_problem_create(model_.handle(), (int)max_priority_level_, &h_), "ikb_problem_create");
        for (const auto &tp : ordered_) {
            const Task &t = *tp.first;
            int rc = 0;
            if (t.kind() == Task::Kind::Frame) {
                const auto &f = static_cast<const FrameTask &>(t);
                rc = ikb_problem_add_frame_task(h_, frame_id(f.frame), (int)f.type, frame_id(f.reference_frame),
                                                (int)tp.second, t.weighting().data());
            } else if (t.kind() == Task::Kind::AlignAxis) {
                const auto &a = static_cast<const AlignAxisTask &>(t);
                rc = ikb_problem_add_align_axis_task(h_, frame_id(a.frame), (int)a.axis, frame_id(a.reference_frame),
                                                     (int)tp.second, t.weighting().data());
            } else if (t.kind() == Task::Kind::CentreOfMass) {
                const auto &c = static_cast<const CentreOfMassTask &>(t);
                rc = ikb_problem_add_com_task(h_, frame_id(c.reference_frame), (int)tp.second, t.weighting().data());
            } else {
                const auto &p = static_cast<const PostureTask &>(t);
                rc = ikb_problem_add_posture_task(h_, (int)p.nj, (int)tp.second, t.weighting().data(), p.mask.data());
            }
            if (rc < 0) {
                const std::string msg = ikb_last_error();
                release();
                throw std::runtime_error("InverseKinematicsProblem: " + msg);
            }
        }
        for (const auto &c : constraints_)
            if (ikb_problem_add_frame_constraint(h_, frame_id(c->frame), (int)c->type, frame_id(c->reference_frame)) < 0) {
                const std::string msg = ikb_last_error();
                release();
                throw std::runtime_error("InverseKinematicsProblem: " + msg);
            }
        check(ikb_problem_finalize(h_, device), "ikb_problem_finalize");
        baked_weights_ = w;
        device_ = device;
        return h_;
    }
    // Per-problem target vector gathered from the tasks' public `target` members (insertion order).
    vector_t gather_targets() const {
        vector_t t;
        for (const auto &tp : ordered_) {
            const Task &k = *tp.first;
            if (k.kind() == Task::Kind::Frame) {
                const auto &f = static_cast<const FrameTask &>(k);
                t.insert(t.end(), f.target.R.begin(), f.target.R.end());
                t.insert(t.end(), f.target.p.begin(), f.target.p.end());
            } else if (k.kind() == Task::Kind::AlignAxis) {
                const auto &a = static_cast<const AlignAxisTask &>(k);
                t.insert(t.end(), a.target.begin(), a.target.end());
            } else if (k.kind() == Task::Kind::CentreOfMass) {
                const auto &c = static_cast<const CentreOfMassTask &>(k);
                t.insert(t.end(), c.target.begin(), c.target.end());
            } else {
                const auto &p = static_cast<const PostureTask &>(k);
                t.insert(t.end(), p.target.begin(), p.target.end());
            }
        }
        return t;
    }

   private:
    template <class Map, class Vec, class Ptr>
    Ptr add(Map &map, Vec &vec, const std::string &name, const Ptr &task, std::size_t priority) {
        if (priority > max_priority_level_) throw std::out_of_range("Maximum priority level exceeded!");
        map.insert({name, vec.size()});
        vec.push_back(task);
        tasks_[priority].push_back(task);
        ordered_.emplace_back(task, priority);
        release();
        return vec.back();
    }
    int frame_id(const std::string &name) const {
        const index_t id = model_.getFrameId(name);
        // the reference indexes out of range here (problem.hpp:85-92, common.hpp:50); this is a hard error instead
        if (id >= (index_t)model_.nframes) throw std::runtime_error("unknown frame '" + name + "'");
        return (int)id;
    }
    void release() {
        if (h_) ikb_problem_free(h_);
        h_ = nullptr;
    }
    model_t model_;
    std::size_t max_priority_level_;
    std::vector<std::vector<std::shared_ptr<Task>>> tasks_;
    std::vector<std::pair<std::shared_ptr<Task>, std::size_t>> ordered_;
    std::vector<std::shared_ptr<PostureTask>> posture_tasks_;
    std::unordered_map<string_t, std::size_t> posture_tasks_map_;
    std::vector<std::shared_ptr<FrameTask>> frame_tasks_;
    std::unordered_map<string_t, std::size_t> frame_tasks_map_;
    std::vector<std::shared_ptr<AlignAxisTask>> axis_tasks_;
    std::unordered_map<string_t, std::size_t> axis_tasks_map_;
    std::shared_ptr<CentreOfMassTask> com_task_ = nullptr;
    std::vector<std::shared_ptr<FrameConstraint>> constraints_;
    std::unordered_map<string_t, std::size_t> constraints_map_;
    ikb_problem *h_ = nullptr;
    vector_t baked_weights_;
    int device_ = 0;
};

// ik/ik/visitor.hpp:7-24.  The stock stop test  ||e[0]||^2 < tolerance  (priority-0 rows, 1e-4 in the reference) runs
// inside the kernel.  A class that OVERRIDES should_stop is honoured by ik::dls / ik::pik: the loop of dls.cpp:14-74 then
// runs on the host, one device iteration per step, and the override sees e and dq where the reference calls it.
class InverseKinematicsProblem;
class inverse_kinematics_visitor {
   public:
    inverse_kinematics_visitor() = default;
    virtual ~inverse_kinematics_visitor() = default;
    number_t tolerance = 1e-4;
    // visitor.hpp:15-21: e = the weighted error vector of every priority level, dq = the step just computed
    virtual bool should_stop(const InverseKinematicsProblem &problem, const std::vector<vector_t> &e, const vector_t &dq) const {
        (void)problem; (void)dq;
        number_t s = 0;
        for (number_t x : e.at(0)) s += x * x;
        return s < tolerance;
    }
    // true for the stock test (which the kernel evaluates); an overriding class returns false
    virtual bool is_default_test() const { return true; }
};
class default_inverse_kinematics_visitor : public inverse_kinematics_visitor {};

struct dls_parameters : public default_solver_parameters {  // ik/ik/dls.hpp:24-28
    number_t damping = 1e-2;
    bool random_restart = false;  // declared by the reference, never read
};

struct dls_info {  // ik/ik/dls.hpp:71-74 (never filled by the reference; filled here)
    bool success = false;
    int iterations = 0;
};

class dls_data {  // ik/ik/dls.hpp:34-65 + ik/ik/data.hpp:8-28: the user-owned per-solve record
   public:
    explicit dls_data(const InverseKinematicsProblem &problem) : q(problem.model().nq, 0.0), dq(problem.model().nv, 0.0) {
        std::size_t rows = 0;
        for (std::size_t l = 0; l <= problem.max_priority_level(); ++l) {
            e.emplace_back(problem.e_size(l), 0.0);                                // data.hpp:24, data.cpp:15-18
            J.emplace_back(problem.e_size(l) * (std::size_t)problem.model().nv, 0.0);
            rows += problem.e_size(l);
        }
        et.assign(rows, 0.0);
        Jt.assign(rows * (std::size_t)problem.model().nv, 0.0);
    }
    bool success = false;
    vector_t q;
    vector_t dq;                 // problem_data::dq: the last step direction computed (dls.cpp:52)
    std::vector<vector_t> e;     // weighted task errors per priority level, as of the last evaluation (data.hpp:24)
    std::vector<vector_t> J;     // weighted task Jacobians per priority level, row-major [e_size(l)][nv] (data.hpp:26)
    vector_t et, Jt;             // the stacked copies ik::dls works on (dls.hpp:54-57, dls.cpp:18-24), Jt row-major [rows][nv]
    dls_info info;
    number_t residual = 0;  // ||e[0]||^2 at the last evaluation (visitor.hpp:19)
#ifdef IK_HAVE_EIGEN
    Eigen::Map<const Eigen::VectorXd> q_eigen() const { return Eigen::Map<const Eigen::VectorXd>(q.data(), (Eigen::Index)q.size()); }
    Eigen::Map<const Eigen::VectorXd> dq_eigen() const { return Eigen::Map<const Eigen::VectorXd>(dq.data(), (Eigen::Index)dq.size()); }
    Eigen::Map<const Eigen::VectorXd> e_eigen(std::size_t level) const { return Eigen::Map<const Eigen::VectorXd>(e.at(level).data(), (Eigen::Index)e.at(level).size()); }
    Eigen::Map<const Eigen::Matrix<number_t, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>> J_eigen(std::size_t level) const {
        const Eigen::Index nv = (Eigen::Index)dq.size();
        return Eigen::Map<const Eigen::Matrix<number_t, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor>>(J.at(level).data(), (Eigen::Index)e.at(level).size(), nv);
    }
#endif
    // (bridge) split the stacked e / J of the last evaluation into the per-level members
    void scatter_levels(std::size_t nv) {
        std::size_t r0 = 0;
        for (std::size_t l = 0; l < e.size(); ++l) {
            const std::size_t m = e[l].size();
            for (std::size_t i = 0; i < m; ++i) e[l][i] = et[r0 + i];
            for (std::size_t i = 0; i < m * nv; ++i) J[l][i] = Jt[r0 * nv + i];
            r0 += m;
        }
    }
};

inline ikb_dls_params to_c(const dls_parameters &p, const inverse_kinematics_visitor &v) {
    ikb_dls_params c;
    ikb_dls_params_default(&c);
    c.max_iterations = (int32_t)p.max_iterations;
    c.max_time = p.max_time;
    c.step_length = p.step_length;
    c.damping = p.damping;
    c.random_restart = p.random_restart ? 1 : 0;
    c.tolerance = v.tolerance;
    return c;
}

// vector_t ik::dls(problem, q0, data, visitor, p) -- reference ik/ik/dls.hpp:111-114, dls.cpp:5-78.  `data` receives what
// the reference leaves in it: success, q, dq, e, J of the last evaluation (data.hpp:15-28).
inline vector_t dls(InverseKinematicsProblem &problem, const vector_t &q0, dls_data &data,
                    const inverse_kinematics_visitor &visitor = inverse_kinematics_visitor(),
                    const dls_parameters &p = dls_parameters()) {
    if ((int)q0.size() != problem.model().nq) throw std::invalid_argument("dls: q0 has the wrong size");
    const vector_t targets = problem.gather_targets();
    const std::size_t nv = (std::size_t)problem.model().nv;
    int ok = 0, it = 0;
    data.q.assign(q0.size(), 0.0);
    data.dq.assign(nv, 0.0);
    if (visitor.is_default_test()) {
        const ikb_dls_params c = to_c(p, visitor);
        check(ikb_dls_solve_ex(problem.handle(), &c, q0.data(), targets.data(), data.q.data(), &ok, &it, &data.residual, data.dq.data(),
                               data.et.data(), data.Jt.data()), "ikb_dls_solve_ex");
        data.scatter_levels(nv);
    } else {
        // overridden stop test: dls.cpp:14-74 on the host, one device iteration (tolerance < 0: never stops itself) per step
        dls_parameters one = p;
        one.max_iterations = 1;
        inverse_kinematics_visitor never;
        never.tolerance = -1.0;
        const ikb_dls_params c = to_c(one, never);
        vector_t q = q0, qn(q0.size());
        for (std::size_t i = 0; i < p.max_iterations && !ok; ++i) {
            int ok1 = 0, it1 = 0;
            check(ikb_dls_solve_ex(problem.handle(), &c, q.data(), targets.data(), qn.data(), &ok1, &it1, &data.residual, data.dq.data(),
                                   data.et.data(), data.Jt.data()), "ikb_dls_solve_ex");
            data.scatter_levels(nv);
            if (visitor.should_stop(problem, data.e, data.dq)) ok = 1;   // dls.cpp:61-64: the un-stepped iterate is returned
            else { q = qn; ++it; }
        }
        data.q = q;
    }
    data.success = ok != 0;
    data.info.success = data.success;
    data.info.iterations = it;
    return data.q;
}
#ifdef IK_HAVE_EIGEN
// the reference's exact signature: vector_t = Eigen::VectorXd (common.hpp:28, dls.hpp:111-114)
inline Eigen::VectorXd dls(InverseKinematicsProblem &problem, const Eigen::Ref<const Eigen::VectorXd> &q0, dls_data &data,
                           const inverse_kinematics_visitor &visitor = inverse_kinematics_visitor(),
                           const dls_parameters &p = dls_parameters()) {
    const vector_t q = dls(problem, vector_t(q0.data(), q0.data() + q0.size()), data, visitor, p);
    return Eigen::Map<const Eigen::VectorXd>(q.data(), (Eigen::Index)q.size());
}
#endif

// Batched extension: B independent problems sharing the task list; host arrays, row-major [B][nq] / [B][target_size]
// (target layout: tasks in insertion order; FrameTask = 9 rotation (row-major) + 3 translation scalars).
struct dls_batch_result {
    std::vector<number_t> q;           // [B][nq]
    std::vector<std::uint8_t> success; // [B]
    std::vector<std::int32_t> iterations;
    std::vector<number_t> residual;
};
inline dls_batch_result dls_batch(InverseKinematicsProblem &problem, std::size_t B, const number_t *q0, const number_t *targets,
                                  const inverse_kinematics_visitor &visitor = inverse_kinematics_visitor(),
                                  const dls_parameters &p = dls_parameters()) {
    ikb_problem *h = problem.handle();
    const int nq = problem.model().nq, tsz = ikb_problem_target_size(h);
    dls_batch_result r;
    r.q.resize(B * nq);
    r.success.resize(B);
    r.iterations.resize(B);
    r.residual.resize(B);
    const ikb_dls_params c = to_c(p, visitor);
    ikb_batch_io io = {};
    io.q0 = q0; io.q0_elem_stride = 1; io.q0_batch_stride = nq;
    io.targets = targets; io.targets_elem_stride = 1; io.targets_batch_stride = tsz;
    io.q = r.q.data(); io.q_elem_stride = 1; io.q_batch_stride = nq;
    io.success = r.success.data(); io.iters = r.iterations.data(); io.resid = r.residual.data();
    check(ikb_dls_solve_batch_host(h, IKB_F64, &c, (int64_t)B, &io), "ikb_dls_solve_batch_host");
    return r;
}

// The same batch sharded over several GPUs (north_star: "the batch shards naturally across the 8 GPUs of one box"):
// contiguous slices, one per listed device, solved concurrently; results land in the returned host arrays (ikb_multi_*).
inline dls_batch_result dls_batch(InverseKinematicsProblem &problem, std::size_t B, const number_t *q0, const number_t *targets,
                                  const std::vector<int> &devices,
                                  const inverse_kinematics_visitor &visitor = inverse_kinematics_visitor(),
                                  const dls_parameters &p = dls_parameters()) {
    ikb_problem *h = problem.handle(devices.empty() ? 0 : devices[0]);
    const int nq = problem.model().nq, tsz = ikb_problem_target_size(h);
    dls_batch_result r;
    r.q.resize(B * nq);
    r.success.resize(B);
    r.iterations.resize(B);
    r.residual.resize(B);
    const ikb_dls_params c = to_c(p, visitor);
    ikb_batch_io io = {};
    io.q0 = q0; io.q0_elem_stride = 1; io.q0_batch_stride = nq;
    io.targets = targets; io.targets_elem_stride = 1; io.targets_batch_stride = tsz;
    io.q = r.q.data(); io.q_elem_stride = 1; io.q_batch_stride = nq;
    io.success = r.success.data(); io.iters = r.iterations.data(); io.resid = r.residual.data();
    ikb_multi *m = nullptr;
    const int rc = ikb_multi_create(h, devices.data(), (int)devices.size(), 2, 1, &m);
    std::unique_ptr<ikb_multi, void (*)(ikb_multi *)> guard(m, ikb_multi_free);
    check(rc, "ikb_multi_create");
    check(ikb_multi_dls_solve_batch_host(m, IKB_F64, &c, (int64_t)B, &io), "ikb_multi_dls_solve_batch_host");
    return r;
}

// ---- ik::pik: priority-based IK (ik/ik/pik.hpp:13-57, pik.cpp:31-96) ----
struct pik_parameters {  // pik.hpp:13-18
    int max_iterations = 100;
    double damping = 1e-2;  // declared by the reference, never read: pik.cpp uses pik_data::lambda
    double step_length = 1.0;
    double max_time = 1.0;
};

class pik_data {  // pik.hpp:27-49 + data.hpp:8-28 (P and da live in the kernel; da is zero in the reference)
   public:
    explicit pik_data(const InverseKinematicsProblem &problem)
        : q(problem.model().nq, 0.0), dq(problem.model().nv, 0.0), lambda(problem.max_priority_level() + 1, 1.0) {}
    bool success = false;
    vector_t q;
    vector_t dq;
    std::vector<double> lambda;  // damping factor of every priority level (pik.hpp:31: 1.0)
    dls_info info;
    number_t residual = 0;
};

// vector_t ik::pik(problem, q0, data, visitor, p) -- reference pik.hpp:51-54
inline vector_t pik(InverseKinematicsProblem &problem, const vector_t &q0, pik_data &data,
                    const inverse_kinematics_visitor &visitor = inverse_kinematics_visitor(),
                    const pik_parameters &p = pik_parameters()) {
    if ((int)q0.size() != problem.model().nq) throw std::invalid_argument("pik: q0 has the wrong size");
    ikb_problem *h = problem.handle();
    const int nq = problem.model().nq, tsz = ikb_problem_target_size(h);
    ikb_pik_params c;
    ikb_pik_params_default(&c);
    c.max_iterations = p.max_iterations;
    c.step_length = p.step_length;
    c.tolerance = visitor.tolerance;
    for (std::size_t l = 0; l < data.lambda.size() && l < 7; ++l) c.lambda[l] = data.lambda[l];
    const vector_t targets = problem.gather_targets();
    int ok = 0, it = 0;
    data.q.assign(q0.size(), 0.0);
    data.dq.assign((std::size_t)problem.model().nv, 0.0);
    (void)nq; (void)tsz;
    check(ikb_pik_solve_ex(h, &c, q0.data(), targets.data(), data.q.data(), &ok, &it, &data.residual, data.dq.data(), nullptr, nullptr),
          "ikb_pik_solve_ex");
    data.success = ok != 0;
    data.info.success = data.success;
    data.info.iterations = it;
    return data.q;
}

// Stream of batches (extension, ikb_queue_*): keeps several batches in flight and launches `merge` consecutive ones as
// one kernel pair; host-buffer copies of one group run beside the kernels of its neighbours.  The batched analogue of the
// reference's per-tick loop (ik_ros/src/cassie.cpp:112-113, 146-171).
//
//     ik::dls_batch_queue queue(problem, /*depth=*/8, /*merge=*/4);
//     auto t = queue.submit(B, q0, targets, result);   // returns at once; `result` must outlive the batch
//     ...                                               // submit more batches
//     queue.wait(t);                                    // `result` is complete
class dls_batch_queue {
   public:
    using ticket_t = std::int64_t;
    dls_batch_queue(InverseKinematicsProblem &problem, int depth = 8, int merge = 4) : problem_(problem) {
        check(ikb_queue_create(problem.handle(), depth, merge, &q_), "ikb_queue_create");
    }
    ~dls_batch_queue() { ikb_queue_free(q_); }
    dls_batch_queue(const dls_batch_queue &) = delete;
    dls_batch_queue &operator=(const dls_batch_queue &) = delete;

    // host arrays as in ik::dls_batch; `out` is resized here and filled when the batch completes
    ticket_t submit(std::size_t B, const number_t *q0, const number_t *targets, dls_batch_result &out,
                    const inverse_kinematics_visitor &visitor = inverse_kinematics_visitor(), const dls_parameters &p = dls_parameters()) {
        ikb_problem *h = problem_.handle();
        const int nq = problem_.model().nq, tsz = ikb_problem_target_size(h);
        out.q.resize(B * nq);
        out.success.resize(B);
        out.iterations.resize(B);
        out.residual.resize(B);
        const ikb_dls_params c = to_c(p, visitor);
        ikb_batch_io io = {};
        io.q0 = q0; io.q0_elem_stride = 1; io.q0_batch_stride = nq;
        io.targets = targets; io.targets_elem_stride = 1; io.targets_batch_stride = tsz;
        io.q = out.q.data(); io.q_elem_stride = 1; io.q_batch_stride = nq;
        io.success = out.success.data(); io.iters = out.iterations.data(); io.resid = out.residual.data();
        const ticket_t t = ikb_queue_submit_host(q_, IKB_F64, &c, (int64_t)B, &io);
        if (t < 0) check((int)-t, "ikb_queue_submit_host");
        return t;
    }
    void wait(ticket_t t) { check(ikb_queue_wait(q_, t), "ikb_queue_wait"); }
    void flush() { check(ikb_queue_flush(q_), "ikb_queue_flush"); }
    void drain() { check(ikb_queue_drain(q_), "ikb_queue_drain"); }

   private:
    InverseKinematicsProblem &problem_;
    ikb_queue *q_ = nullptr;
};

}  // namespace ik
