// test_host.cpp -- the C++ host layer exercised the way the reference's own tests
// exercise the Rust API: state.rs:166-185 (euler_step) and state.rs:218-280 (rk4_step)
// verbatim, plus the flocking::Simulation call sequence of demos/flocking.rs:105-121,215-227.
//   test_host          host-only checks (no GPU needed)
//   test_host --gpu    everything
#include <cmath>
#include <cstdio>
#include <cstring>

#include "feriphys_cuda.hpp"

using namespace feriphys;
using namespace feriphys::simulation;

#define CHECK(c)                                                         \
    do {                                                                 \
        if (!(c)) {                                                      \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                    \
        }                                                                \
    } while (0)

// state.rs:120-164
struct Point {
    Vector3 position, velocity;
    static size_t num_state_elements() { return 6; }
    static Point from_state_vector(std::vector<float> d) {
        if (d.size() != num_state_elements()) throw Panic("State Vector incorrect size!");
        return {{d[0], d[1], d[2]}, {d[3], d[4], d[5]}};
    }
    std::vector<float> derivative() const { return {velocity[0], velocity[1], velocity[2], 1.0f, -1.0f, 0.0f}; }
    std::vector<float> as_state() const {
        return {position[0], position[1], position[2], velocity[0], velocity[1], velocity[2]};
    }
};
// state.rs:187-216
struct ExampleFn {
    float y, t, timestep;
    static size_t num_state_elements() { return 3; }
    static ExampleFn from_state_vector(std::vector<float> d) {
        if (d.size() != num_state_elements()) throw Panic("State Vector incorrect size!");
        return {d[0], d[1], d[2]};
    }
    std::vector<float> derivative() const { return {y - t * t + 1.0f, 1.0f, 0.0f}; }
    std::vector<float> as_state() const { return {y, t, timestep}; }
};

// springy_mesh.rs:181-257 with the device kernel named: the whole step runs on the GPU
struct SpringyPoint {
    float mass;
    Vector3 position, velocity, accumulated_force;
    static size_t num_state_elements() { return 10; }
    static int device_kind() { return FP_STATEFUL_SPRINGY_POINT; }
    static SpringyPoint from_state_vector(std::vector<float> d) {
        if (d.size() != num_state_elements()) throw Panic("State Vector incorrect size!");
        return {d[0], {d[1], d[2], d[3]}, {d[4], d[5], d[6]}, {d[7], d[8], d[9]}};
    }
    std::vector<float> derivative() const {
        return {0.0f, velocity[0], velocity[1], velocity[2], accumulated_force[0] / mass, accumulated_force[1] / mass,
                accumulated_force[2] / mass, 0.0f, 0.0f, 0.0f};
    }
    std::vector<float> as_state() const {
        return {mass, position[0], position[1], position[2], velocity[0], velocity[1], velocity[2],
                accumulated_force[0], accumulated_force[1], accumulated_force[2]};
    }
};
// the same type WITHOUT the addition: derivative() on the host, vector arithmetic on the device
struct SpringyPointHost : SpringyPoint {
    static SpringyPointHost from_state_vector(std::vector<float> d) { return {SpringyPoint::from_state_vector(std::move(d))}; }
};

static int host_only() {
    CHECK(flocking::Config().dt == 0.001f);
    CHECK(Duration::from_secs_f32(2.7f).secs == 2 && Duration::from_secs_f32(2.7f).nanos == 700000048u);
    CHECK(Duration::from_secs_f32(-0.0f).is_zero());
    bool threw = false;
    try { Duration::from_secs_f32(-1.0f); } catch (const Panic &) { threw = true; }
    CHECK(threw);
    CHECK(Duration::from_secs(4).as_secs_f32() == 4.0f);
    // boid.rs:46-53: returns path(t) then advances; first step has zero velocity (F9)
    flocking::LeadBoid lead([](float t) { return Vector3{25.0f * std::cos(t / 12.0f), 0.5f, 0.0f}; });
    CHECK(lead.position()[0] == 25.0f && lead.weight() == 10.0f);
    lead.step(Duration::from_secs_f32(0.001f));
    CHECK(lead.position()[0] == 25.0f && lead.velocity()[0] == 0.0f);
    lead.step(Duration::from_secs_f32(0.001f));
    CHECK(lead.position()[0] == 25.0f * std::cos(0.001f / 12.0f));
    // a wrong-size chunk panics like the reference's Stateful impls (springy_mesh.rs:205-207)
    threw = false;
    try { state::State<Point>::from_state_vector({1, 2, 3, 4, 5, 6, 7}); } catch (const Panic &) { threw = true; }
    CHECK(threw);
    std::puts("CPP_HOST_OK");
    return 0;
}

static int with_gpu() {
    {   // state.rs:166-185
        const float h = 0.5f;
        state::State<Point> st({Point{{0, 0, 0}, {0, 0, 1}}});
        CHECK((st.as_vector() == std::vector<float>{0, 0, 0, 0, 0, 1}));
        auto pts = st.euler_step(h).get_elements();
        CHECK((pts[0].position == Vector3{0.0f, 0.0f, 0.5f}));
        CHECK((pts[0].velocity == Vector3{0.5f, -0.5f, 1.0f}));
    }
    {   // state.rs:218-280
        const float h = 0.5f, acceptable_error = 0.005f;
        const double golden[4] = {1.425130208333333, 2.640859085770477, 4.009155464830968, 5.305471950534675};
        const float ts[4] = {0.5f, 1.0f, 1.5f, 2.0f};
        std::vector<ExampleFn> v{{0.5f, 0.0f, h}};
        for (int k = 0; k < 4; ++k) {
            v = state::State<ExampleFn>(std::move(v)).rk4_step(h).get_elements();
            CHECK(golden[k] + acceptable_error > v[0].y && golden[k] - acceptable_error < v[0].y);
            CHECK(v[0].t == ts[k] && v[0].timestep == 0.5f);
        }
    }
    {   // State<Point> of the springy mesh: device-resident kernel == host derivative + device combine
        std::vector<SpringyPoint> dev;
        std::vector<SpringyPointHost> host;
        for (int i = 0; i < 1000; ++i) {
            const float f = (float)i;
            SpringyPoint p{1.0f + 0.01f * f, {f, 2.0f * f, -f}, {0.5f, -0.25f * f, 0.125f}, {0.3f * f, -9.8f, 0.7f}};
            dev.push_back(p);
            host.push_back({p});
        }
        auto a = state::State<SpringyPoint>(dev).rk4_step(0.01f).euler_step(0.02f).as_vector();
        auto b = state::State<SpringyPointHost>(host).rk4_step(0.01f).euler_step(0.02f).as_vector();
        CHECK(a.size() == 10000 && std::memcmp(a.data(), b.data(), a.size() * 4) == 0);
        CHECK(a[14] != 0.5f);  // (element 1: force 0.3 on mass 1.01 moved its velocity; element 0 carries none)
    }
    {   // the demo's sim 1 (demos/flocking.rs:105-121) stepped headless two ways: must agree bit for bit
        std::vector<float> st;
        uint64_t x = 88172645463325252ull;
        auto u01 = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return (float)(x >> 40) * 0x1p-24f; };
        for (int i = 0; i < 110; ++i) {
            const float s[6] = {25.0f + u01(), 0.5f + u01(), 0.0f + u01(), u01(), u01(), u01()};
            st.insert(st.end(), s, s + 6);
        }
        auto path = [](float t) { return Vector3{25.0f * std::cos(t / 12.0f), 0.5f, 0.0f}; };
        auto make = [&]() {
            return new flocking::Simulation(st, std::nullopt, std::vector<flocking::LeadBoid>{flocking::LeadBoid(path)},
                                            std::vector<flocking::Obstacle>{{{-5.0f, 0.0f, 0.0f}, 4.0f}}, std::nullopt);
        };
        flocking::Simulation *a = make(), *b = make();
        for (int k = 0; k < 100; ++k) CHECK(a->step() == Duration::from_millis(1));
        b->step_many(100);
        const auto sa = a->read_state(), sb = b->read_state();
        CHECK(sa.size() == 660 && std::memcmp(sa.data(), sb.data(), sa.size() * 4) == 0);
        CHECK(std::memcmp(sa.data(), st.data(), sa.size() * 4) != 0);
        CHECK(a->status() == 0);
        const auto inst = a->get_boid_instances();
        CHECK(inst.size() == 110 && inst[0].scale == 0.1f && inst[3].position[0] == sa[18]);
        const float q = inst[7].rotation[0] * inst[7].rotation[0] + inst[7].rotation[1] * inst[7].rotation[1] +
                        inst[7].rotation[2] * inst[7].rotation[2] + inst[7].rotation[3] * inst[7].rotation[3];
        CHECK(std::fabs(q - 1.0f) < 1e-5f);
        CHECK((*a->lead_boids())[0].position() == (*b->lead_boids())[0].position());
        delete a;
        delete b;
    }
    std::puts("CPP_GPU_OK");
    return 0;
}

int main(int argc, char **argv) {
    try {
        if (int rc = host_only()) return rc;
        if (argc > 1 && !std::strcmp(argv[1], "--gpu")) return with_gpu();
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 2;
    }
}
