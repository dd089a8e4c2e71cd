// fp_walk_stage.cuh -- building blocks of the shared-memory staged walk kernels: the TMA
// bulk-copy / mbarrier primitives (fp_walk.cu, fp_walk_nl.cu) and the drain, i.e. the exact force
// phase over a per-thread survivor list (fp_walk_nl.cu).  The production kernel in fp_walk.cu
// still carries its own inlined copy of the drain: folding it onto drain_list changes ptxas'
// register allocation of that kernel (tools/sass_pins.py), so the merge waits until it can be
// timed on a B200.
#pragma once

#include "fp_grid.cuh"

namespace fp {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// The drain of a staged walk: exact forces over the first `nb` entries of a thread's survivor
// list, accumulated into `acc` in list (= slot) order.  An entry is (row << 12 | tile offset);
// `tslot[row]` turns the offset back into a slot for the velocity gather; entry k lives at
// lst[k * BLOCK].  Two entries per trip: their (branch-free) force evaluations are independent
// and interleave; the two adds into acc stay in list order.  Velocities are gathered TWO trips
// ahead (four loads in flight): one trip of arithmetic does not cover an L2 miss (ncu: the
// single-trip version spent 17 % of the kernel in long-scoreboard stalls here).
template <int BLOCK>
__device__ __forceinline__ void drain_list(const DevParams &P, const Self &self, const uint16_t *lst, const int nb,
                                           const float *tx, const float *ty, const float *tz,
                                           const uint32_t *tslot, const uint32_t t_self,
                                           const float4 *__restrict__ vel_s, V3 &acc) {
    auto slot_of = [&](uint32_t e) { return (e & 0xfffu) + tslot[e >> 12]; };
    uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0;
    float4 w0 = make_float4(0, 0, 0, 0), w1 = w0, w2 = w0, w3 = w0;
    if (nb > 0) { e0 = lst[0]; w0 = __ldg(vel_s + slot_of(e0)); }
    if (nb > 1) { e1 = lst[BLOCK]; w1 = __ldg(vel_s + slot_of(e1)); }
    if (nb > 2) { e2 = lst[2 * BLOCK]; w2 = __ldg(vel_s + slot_of(e2)); }
    if (nb > 3) { e3 = lst[3 * BLOCK]; w3 = __ldg(vel_s + slot_of(e3)); }
    for (int k = 0; k < nb; k += 2) {
        const uint32_t ta = e0, tb = e1;
        const float4 va = w0, vb = w1;
        const bool hasb = k + 1 < nb;
        e0 = e2; e1 = e3; w0 = w2; w1 = w3;
        if (k + 4 < nb) { e2 = lst[(k + 4) * BLOCK]; w2 = __ldg(vel_s + slot_of(e2)); }
        if (k + 5 < nb) { e3 = lst[(k + 5) * BLOCK]; w3 = __ldg(vel_s + slot_of(e3)); }
        const uint32_t ia = ta & 0xfffu, ib = tb & 0xfffu;
        const V3 pa = v3(tx[ia], ty[ia], tz[ia]), pb = v3(tx[ib], ty[ib], tz[ib]);
        V3 da, db;
        const float ma = pair_m2(self, v3(pa.x, pa.y, pa.z), da);
        const float mb = pair_m2(self, v3(pb.x, pb.y, pb.z), db);
        // what counts: not the boid itself (flocking.rs:137-139), and in range by the
        // EXACT squared distance (the pre-gate let a sliver too many through)
        const bool oka = ia != t_self && !(ma >= P.m2_cut);
        const bool okb = hasb && ib != t_self && !(mb >= P.m2_cut);
        const bool fast = P.fast_ok && (!oka || (ma >= FAST_M2_LO && ma <= FAST_M2_HI)) &&
                          (!okb || (mb >= FAST_M2_LO && mb <= FAST_M2_HI));
        if (fast) {  // (a lane that does not count may hold inf / NaN; it is never used)
            bool visa, visb;
            const V3 fa = pair_force_fast(P, self, da, ma, v3(va.x, va.y, va.z), visa);
            const V3 fb = pair_force_fast(P, self, db, mb, v3(vb.x, vb.y, vb.z), visb);
            if (oka && visa) acc = vadd(acc, fa);
            if (okb && visb) acc = vadd(acc, fb);
        } else {  // extreme distances (coincident boids, ...): generic exact path
            V3 contrib;
            if (oka && pair_inrange<false>(P, self, da, ma, v3(va.x, va.y, va.z), 1.0f, P.cstar, contrib))
                acc = vadd(acc, contrib);
            if (okb && pair_inrange<false>(P, self, db, mb, v3(vb.x, vb.y, vb.z), 1.0f, P.cstar, contrib))
                acc = vadd(acc, contrib);
        }
    }
}

}  // namespace fp
