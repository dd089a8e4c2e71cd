// feriphys_cuda.hpp -- C++17 host layer over the C ABI (include/feriphys_cuda.h).
//
// Mirrors the reference's Rust API for the flocking path -- same names, argument
// meaning and error behaviour -- so code written against
//   feriphys::simulation::flocking::{Config, Simulation, boid::LeadBoid, obstacle::Obstacle},
//   simulation::{parametric::Parametric, point_attractor::PointAttractor,
//                bounding_box::BoundingBox, state::{Stateful, State, Integration}}
// reads the same here.  The reference is compiled code, so this layer is C++; the
// Rust crate that binds the same C ABI is rust/feriphys-cuda (not buildable in this
// image).  All arithmetic on boids happens in libferiphys_cuda.so; this header only
// owns what cannot cross an FFI: the lead boids' path functions (boid.rs:35,
// parametric.rs:5), evaluated here in the reference's order.
//
// Where Rust panics this layer throws feriphys::Panic.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/feriphys_cuda.h"

namespace feriphys {

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

using Vector3 = std::array<float, 3>;

inline void fp_check(int rc) {
    if (rc != FP_OK) throw std::runtime_error(std::string("feriphys-cuda: ") + fp_last_error());
}

// std::time::Duration: whole seconds + nanoseconds.
struct Duration {
    uint64_t secs = 0;
    uint32_t nanos = 0;
    static Duration from_secs(uint64_t s) { return {s, 0}; }
    static Duration from_millis(uint64_t ms) { return {ms / 1000, (uint32_t)(ms % 1000) * 1000000u}; }
    // exact value * 1e9 rounded to nearest-even ns; panics on negative / NaN / overflow
    static Duration from_secs_f32(float x) {
        if (x < 0.0f) throw Panic("can not convert float seconds to Duration: value is negative");
        if (!(x < 18446744073709551616.0f))
            throw Panic("can not convert float seconds to Duration: value is either too big or NaN");
        if (x >= 8388608.0f) return {(uint64_t)x, 0};
        const uint64_t ns = (uint64_t)std::nearbyint((double)x * 1e9);  // product exact in binary64
        return {ns / 1000000000ull, (uint32_t)(ns % 1000000000ull)};
    }
    float as_secs_f32() const { return (float)secs + (float)nanos / 1000000000.0f; }
    bool is_zero() const { return secs == 0 && nanos == 0; }
    bool operator<(const Duration &o) const { return secs < o.secs || (secs == o.secs && nanos < o.nanos); }
    bool operator==(const Duration &o) const { return secs == o.secs && nanos == o.nanos; }
    // impl Add / Sub for Duration: carry in nanoseconds; Sub panics on underflow
    Duration operator+(const Duration &o) const {
        Duration r{secs + o.secs, nanos + o.nanos};
        if (r.nanos >= 1000000000u) {
            r.nanos -= 1000000000u;
            ++r.secs;
        }
        return r;
    }
    Duration operator-(const Duration &o) const {
        if (*this < o) throw Panic("overflow when subtracting durations");
        Duration r{secs - o.secs, nanos};
        if (r.nanos < o.nanos) {
            --r.secs;
            r.nanos += 1000000000u;
        }
        r.nanos -= o.nanos;
        return r;
    }
};

namespace simulation {

// parametric.rs:4-22
class Parametric {
  public:
    using Path = std::function<Vector3(float)>;
    explicit Parametric(Path path) : path_(std::move(path)) {}
    Vector3 step(float dt) {  // returns path(t) THEN advances t
        const Vector3 p = path_(curr_time_);
        curr_time_ = curr_time_ + dt;
        return p;
    }
    float curr_time() const { return curr_time_; }

  private:
    Path path_;
    float curr_time_ = 0.0f;
};

// point_attractor.rs:9-12, bounding_box.rs:5-9
struct PointAttractor {
    Vector3 position;
    float mass;
};
struct BoundingBox {
    std::pair<float, float> x_range, y_range, z_range;  // Range<f32>: (start, end)
};

namespace flocking {

// graphics/instance.rs:7-11 (rotation: quaternion s, x, y, z)
struct Instance {
    Vector3 position;
    std::array<float, 4> rotation;
    float scale;
};

// boid.rs:13-53
class LeadBoid {
  public:
    explicit LeadBoid(Parametric::Path path) : parametric_(path), position_(path(0.0f)) {}
    static LeadBoid make(Parametric::Path path) { return LeadBoid(std::move(path)); }  // LeadBoid::new
    Vector3 position() const { return position_; }
    Vector3 velocity() const { return velocity_; }
    float weight() const { return weight_; }
    void step(Duration dt) {
        if (dt.is_zero()) return;
        const float s = dt.as_secs_f32();
        const Vector3 np = parametric_.step(s);
        for (int a = 0; a < 3; ++a) velocity_[a] = (np[a] - position_[a]) / s;
        position_ = np;
    }

  private:
    Parametric parametric_;
    Vector3 position_;
    Vector3 velocity_{0.0f, 0.0f, 0.0f};
    float weight_ = 10.0f;
};

// obstacle.rs:11-14
struct Obstacle {
    Vector3 position;
    float radius;
    // Obstacle::from_entity (obstacle.rs:48-59): radius = instance.scale * radius
    static std::vector<Obstacle> from_entity(const std::vector<Instance> &instances, float radius) {
        std::vector<Obstacle> out;
        for (const auto &i : instances) out.push_back({i.position, i.scale * radius});
        return out;
    }
};

// flocking.rs:15-51
struct Config {
    float dt = Duration::from_millis(1).as_secs_f32();
    float avoidance_factor = 1.0f;
    float centering_factor = 0.1f;
    float velocity_matching_factor = 0.5f;
    float distance_weight_threshold = 15.0f;
    float distance_weight_threshold_falloff = 1.0f;
    float max_sight_angle = 3.14159274101257324f / 2.0f;
    float max_sight_angle_to_lead_boid = 3.14159274101257324f;
    Duration time_to_start_steering = Duration::from_secs(4);
    bool steering_overrides = false;

    fp_config to_c() const {
        fp_config c{};
        c.dt = dt;
        c.avoidance_factor = avoidance_factor;
        c.centering_factor = centering_factor;
        c.velocity_matching_factor = velocity_matching_factor;
        c.distance_weight_threshold = distance_weight_threshold;
        c.distance_weight_threshold_falloff = distance_weight_threshold_falloff;
        c.max_sight_angle = max_sight_angle;
        c.max_sight_angle_to_lead_boid = max_sight_angle_to_lead_boid;
        c.time_to_start_steering_secs = time_to_start_steering.secs;
        c.time_to_start_steering_nanos = time_to_start_steering.nanos;
        c.steering_overrides = steering_overrides ? 1 : 0;
        return c;
    }
};

// flocking.rs:53-246
class Simulation {
  public:
    // ADDITION: explicit initial state, n x [px py pz vx vy vz].  Simulation::new's
    // unseeded jitter (flocking.rs:77-82) is the caller's to draw.
    Simulation(const std::vector<float> &state_aos6, std::optional<BoundingBox> bounding_box,
               std::optional<std::vector<LeadBoid>> lead_boids, std::optional<std::vector<Obstacle>> obstacles,
               std::optional<std::vector<PointAttractor>> attractors, int device = 0)
        : lead_boids_(std::move(lead_boids)) {
        if (state_aos6.size() % 6) throw Panic("State Vector incorrect size!");
        n_ = state_aos6.size() / 6;
        const fp_config c = config_.to_c();
        fp_check(fp_flock_create(&h_, &c, n_, state_aos6.data(), device));
        if (bounding_box) {
            const float b[6] = {bounding_box->x_range.first, bounding_box->x_range.second,
                                bounding_box->y_range.first, bounding_box->y_range.second,
                                bounding_box->z_range.first, bounding_box->z_range.second};
            fp_check(fp_flock_set_bbox(h_, b));
        }
        if (obstacles) {
            std::vector<float> o;
            for (const auto &x : *obstacles) o.insert(o.end(), {x.position[0], x.position[1], x.position[2], x.radius});
            fp_check(fp_flock_set_obstacles(h_, (uint32_t)obstacles->size(), o.data()));
        }
        if (attractors) {
            std::vector<float> a;
            for (const auto &x : *attractors) a.insert(a.end(), {x.position[0], x.position[1], x.position[2], x.mass});
            fp_check(fp_flock_set_attractors(h_, (uint32_t)attractors->size(), a.data()));
        }
        push_leads();
    }
    Simulation(const Simulation &) = delete;
    Simulation &operator=(const Simulation &) = delete;
    ~Simulation() { fp_flock_destroy(h_); }

    // Simulation::step (flocking.rs:97-131)
    Duration step() {
        if (lead_boids_) push_leads();
        fp_check(fp_flock_step(h_, 1));
        if (lead_boids_)
            for (auto &l : *lead_boids_) l.step(Duration::from_secs_f32(config_.dt));
        return get_timestep();
    }
    // ADDITION: n steps in one library call, lead rows tabulated exactly as n step() calls would
    Duration step_many(uint32_t n) {
        if (!n) return {};
        if (lead_boids_ && !lead_boids_->empty()) {
            std::vector<float> table;
            for (uint32_t s = 0; s < n; ++s) {
                append_lead_rows(table);
                for (auto &l : *lead_boids_) l.step(Duration::from_secs_f32(config_.dt));
            }
            fp_check(fp_flock_set_lead_table(h_, n, (uint32_t)lead_boids_->size(), table.data()));
        }
        fp_check(fp_flock_step(h_, n));
        if (lead_boids_) push_leads();
        return get_timestep();
    }
    Duration get_timestep() const { return Duration::from_secs_f32(config_.dt); }
    // sync_sim_config_from_ui (flocking.rs:215-228): the UI holds a Config copy
    void sync_sim_config(const Config &ui_state) {
        config_ = ui_state;
        const fp_config c = config_.to_c();
        fp_check(fp_flock_set_config(h_, &c));
    }
    // get_boid_instances (flocking.rs:230-245)
    std::vector<Instance> get_boid_instances() const {
        std::vector<float> raw(n_ * 8);
        fp_check(fp_flock_read_instances(h_, raw.data()));
        std::vector<Instance> out(n_);
        for (size_t i = 0; i < n_; ++i) {
            const float *r = &raw[8 * i];
            out[i] = {{r[0], r[1], r[2]}, {r[3], r[4], r[5], r[6]}, r[7]};
        }
        return out;
    }
    // ADDITIONS (F4)
    std::vector<float> read_state() const {
        std::vector<float> s(n_ * 6);
        fp_check(fp_flock_read_state(h_, s.data()));
        return s;
    }
    uint32_t status() const {
        uint32_t f = 0;
        fp_check(fp_flock_status(h_, &f));
        return f;
    }
    void set_method(int m) { fp_check(fp_flock_set_method(h_, m)); }
    // lazy re-binning of the grid path (ADDITION): skin < 0 = sized from the flock's speed
    void set_rebin(float skin = -1.0f, float plan_scale = 1.0f) { fp_check(fp_flock_set_rebin(h_, skin, plan_scale)); }
    struct RebinInfo { float skin; uint64_t grid_steps, rebins, replayed; };
    RebinInfo rebin_info() {
        RebinInfo r{};
        fp_check(fp_flock_rebin_info(h_, &r.skin, &r.grid_steps, &r.rebins, &r.replayed));
        return r;
    }
    size_t len() const { return n_; }
    fp_flock *handle() const { return h_; }
    const std::optional<std::vector<LeadBoid>> &lead_boids() const { return lead_boids_; }

  private:
    void append_lead_rows(std::vector<float> &t) const {
        for (const auto &l : *lead_boids_) {
            const Vector3 p = l.position(), v = l.velocity();
            t.insert(t.end(), {p[0], p[1], p[2], v[0], v[1], v[2], l.weight()});
        }
    }
    void push_leads() {
        std::vector<float> t;
        if (lead_boids_) append_lead_rows(t);
        fp_check(fp_flock_set_leads(h_, lead_boids_ ? (uint32_t)lead_boids_->size() : 0, t.empty() ? nullptr : t.data()));
    }
    Config config_;
    std::optional<std::vector<LeadBoid>> lead_boids_;
    fp_flock *h_ = nullptr;
    size_t n_ = 0;
};

}  // namespace flocking

// state.rs:4-113.  T models Stateful:
//   static size_t num_state_elements(); static T from_state_vector(std::vector<float>);
//   std::vector<float> derivative() const; std::vector<float> as_state() const;
namespace state {

enum class Integration { Euler, Rk4 };

template <class T>
class State {
  public:
    explicit State(std::vector<T> elements, int device = 0) : elements_(std::move(elements)), device_(device) {}
    static State from_state_vector(const std::vector<float> &v, int device = 0) {
        const size_t k = T::num_state_elements();
        std::vector<T> e;
        for (size_t i = 0; i < v.size(); i += k) {
            const size_t end = std::min(i + k, v.size());  // a short last chunk reaches T, as itertools::chunks does
            e.push_back(T::from_state_vector(std::vector<float>(v.begin() + i, v.begin() + end)));
        }
        return State(std::move(e), device);
    }
    std::vector<float> derivative() const { return flatten([](const T &t) { return t.derivative(); }); }
    std::vector<float> as_vector() const { return flatten([](const T &t) { return t.as_state(); }); }
    // S + S' * h (state.rs:75-83).  A type that names the device kernel evaluating its derivative
    // (`static int device_kind()` -> FP_STATEFUL_*, an ADDITION to the Stateful concept) is stepped
    // entirely on the device -- one streaming pass, fp_state.cu; any other type keeps derivative()
    // on the host and only the vector arithmetic runs on the device.
    State euler_step(float h) const {
        if (const int kind = kind_of<T>(0)) return device_step(kind, h, false);
        const auto s = as_vector(), d = derivative();
        std::vector<float> out(s.size());
        fp_check(fp_state_euler_combine(device_, s.size(), s.data(), d.data(), h, out.data()));
        return from_state_vector(out, device_);
    }
    // classic RK4 (state.rs:86-106): stages evaluated through T::derivative on the host
    State rk4_step(float h) const {
        if (const int kind = kind_of<T>(0)) return device_step(kind, h, true);
        const auto s = as_vector();
        const auto k1 = derivative();
        const auto k2 = from_state_vector(axpy(s, k1, h * 0.5f), device_).derivative();
        const auto k3 = from_state_vector(axpy(s, k2, h * 0.5f), device_).derivative();
        const auto k4 = from_state_vector(axpy(s, k3, h), device_).derivative();
        std::vector<float> out(s.size());
        fp_check(fp_state_rk4_combine(device_, s.size(), s.data(), k1.data(), k2.data(), k3.data(), k4.data(), h,
                                      out.data()));
        return from_state_vector(out, device_);
    }
    std::vector<T> get_elements() && { return std::move(elements_); }
    const std::vector<T> &elements() const { return elements_; }

  private:
    template <class U>
    static auto kind_of(int) -> decltype(U::device_kind()) { return U::device_kind(); }
    template <class U>
    static int kind_of(...) { return 0; }
    State device_step(int kind, float h, bool rk4) const {
        const auto s = as_vector();
        fp_state *st = nullptr;
        fp_check(fp_state_create(&st, device_, kind, elements_.size(), s.data()));
        const int rc = rk4 ? fp_state_rk4_step(st, h, 1) : fp_state_euler_step(st, h, 1);
        std::vector<float> out(s.size());
        const int rc2 = rc ? rc : fp_state_read(st, out.data());
        fp_state_destroy(st);
        fp_check(rc2);
        return from_state_vector(out, device_);
    }
    template <class F>
    std::vector<float> flatten(F f) const {
        std::vector<float> out;
        for (const auto &e : elements_) {
            const auto v = f(e);
            out.insert(out.end(), v.begin(), v.end());
        }
        return out;
    }
    // utils::vec_add(s, utils::scale(k, c)) (utils.rs:5-21): s + k * c, one device pass
    std::vector<float> axpy(const std::vector<float> &s, const std::vector<float> &k, float c) const {
        if (s.size() != k.size()) throw Panic("Cannot multiply vectors of different lengths!");
        std::vector<float> out(s.size());
        fp_check(fp_state_euler_combine(device_, s.size(), s.data(), k.data(), c, out.data()));
        return out;
    }
    std::vector<T> elements_;
    int device_;
};

}  // namespace state
}  // namespace simulation
}  // namespace feriphys
