// demo_flocking.cpp -- headless driver of the reference's flocking demo.
//
// Builds the scene of demos/flocking.rs:92-156 (ship obstacle at (-5,0,0) r = 4; simulation 1:
// 110 boids spawned at (25,0.5,0) behind one lead boid; simulation 2: 55 + 55 boids at
// (15,10,0) and (25,0.5,0) behind two lead boids) and runs its update loop
// (demos/flocking.rs:209-231): per frame, step each simulation while the time accumulator
// holds a timestep, then get_boid_instances() for both -- with the renderer replaced by a
// checksum of what would have been uploaded.  The frame time is fixed (no wall clock in the
// loop), so the output is reproducible.
//
//   demo_flocking [frames=60] [frame_ms=16] [seed=0xFE21F]
//
// Simulation::new's jitter comes from the unseeded thread RNG in the reference
// (flocking.rs:77-82); here it is the keyed splitmix64 generator the tests use
// (feriphys_b200/synth.py), so the C++ and Python drivers produce identical flocks.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "feriphys_cuda.hpp"

using namespace feriphys;
using namespace feriphys::simulation;
using flocking::Instance;
using flocking::LeadBoid;
using flocking::Obstacle;
using flocking::Simulation;

static uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// u01 of synth.py: keyed by (seed, boid index, component), 24-bit mantissa
static float u01(uint64_t seed, uint64_t index, uint64_t component) {
    const uint64_t key = splitmix64(seed) ^ (index * 6 + component);
    return (float)(splitmix64(key) >> 40) * 5.9604644775390625e-08f;
}

// Simulation::new (flocking.rs:63-95) with the jitter made explicit
static std::vector<float> spawn_flock(const std::vector<Vector3> &spawn, size_t num_boids, uint64_t seed) {
    const size_t per = num_boids / spawn.size();  // integer division drops the remainder (flocking.rs:76)
    std::vector<float> s(per * spawn.size() * 6);
    for (size_t k = 0; k < spawn.size(); ++k)
        for (size_t b = 0; b < per; ++b) {
            const size_t i = k * per + b;
            for (int a = 0; a < 3; ++a) {
                s[6 * i + a] = spawn[k][a] + u01(seed, i, a);
                s[6 * i + 3 + a] = u01(seed, i, 3 + a);
            }
        }
    return s;
}

static uint64_t checksum(const std::vector<Instance> &v, uint64_t h) {
    for (const auto &i : v) {
        uint32_t w[8];
        std::memcpy(w, i.position.data(), 12);
        std::memcpy(w + 3, i.rotation.data(), 16);
        std::memcpy(w + 7, &i.scale, 4);
        for (uint32_t x : w) h = splitmix64(h ^ x);
    }
    return h;
}

int main(int argc, char **argv) {
    const int frames = argc > 1 ? std::atoi(argv[1]) : 60;
    const int frame_ms = argc > 2 ? std::atoi(argv[2]) : 16;
    const uint64_t seed = argc > 3 ? std::strtoull(argv[3], nullptr, 0) : 0xFE21Full;
    try {
        // demos/flocking.rs:92-104
        const std::vector<Instance> ship = {{{-5.0f, 0.0f, 0.0f}, {1.0f, 0.0f, 0.0f, 0.0f}, 1.0f}};
        const std::vector<Obstacle> obstacles = Obstacle::from_entity(ship, 4.0f);
        // :107-121
        std::vector<LeadBoid> leads1;
        leads1.push_back(LeadBoid::make([](float t) { return Vector3{25.0f * std::cos(t / 12.0f), 0.5f, 0.0f}; }));
        Simulation sim1(spawn_flock({{25.0f, 0.5f, 0.0f}}, 110, seed), std::nullopt, leads1, obstacles, std::nullopt);
        // :135-156
        std::vector<LeadBoid> leads2;
        leads2.push_back(LeadBoid::make([](float t) {
            return Vector3{15.0f * std::cos(t / 12.0f), 6.0f + 5.0f * std::cos(t / 12.0f), 15.0f * std::sin(t / 12.0f)};
        }));
        leads2.push_back(LeadBoid::make(
            [](float t) { return Vector3{25.0f * std::cos(t / 10.0f), 1.0f, 10.0f * std::sin(t / 9.0f)}; }));
        Simulation sim2(spawn_flock({{15.0f, 10.0f, 0.0f}, {25.0f, 0.5f, 0.0f}}, 110, seed), std::nullopt, leads2,
                        obstacles, std::nullopt);

        // update(), demos/flocking.rs:209-231, with a fixed frame time
        const Duration frame_time{0, (uint32_t)frame_ms * 1000000u};
        Duration acc1{}, acc2{};
        uint64_t h = 0, steps = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; ++f) {
            acc1 = acc1 + frame_time;
            acc2 = acc2 + frame_time;
            while (!(acc1 < sim1.get_timestep())) {
                acc1 = acc1 - sim1.step();
                ++steps;
            }
            while (!(acc2 < sim2.get_timestep())) {
                acc2 = acc2 - sim2.step();
                ++steps;
            }
            h = checksum(sim1.get_boid_instances(), h);
            h = checksum(sim2.get_boid_instances(), h);
        }
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("frames %d steps %llu checksum %016llx status %u/%u wall_ms_per_frame %.3f\n", frames,
                    (unsigned long long)steps, (unsigned long long)h, sim1.status(), sim2.status(),
                    1e3 * secs / (frames ? frames : 1));
    } catch (const std::exception &e) {
        std::fprintf(stderr, "demo_flocking: %s\n", e.what());
        return 1;
    }
    return 0;
}
