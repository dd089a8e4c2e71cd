//! `extern "C"` block for include/feriphys_cuda.h.  One declaration per entry point,
//! same order as the header; see the header for the reference item each one replaces.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const FP_OK: c_int = 0;
pub const FP_METHOD_AUTO: c_int = 0;
pub const FP_METHOD_ALLPAIRS: c_int = 1;
pub const FP_METHOD_GRID: c_int = 2;
pub const FP_METHOD_SMALL: c_int = 3;
pub const FP_NUMERICS_EXACT: c_int = 0;
pub const FP_NUMERICS_FAST: c_int = 1;
pub const FP_STATEFUL_TEST_POINT: c_int = 1;
pub const FP_STATEFUL_TEST_EXAMPLEFN: c_int = 2;
pub const FP_STATEFUL_SPRINGY_POINT: c_int = 3;
pub const FP_STATEFUL_RIGIDBODY: c_int = 4;
pub const FP_STATEFUL_BOID: c_int = 5;

/// flocking::Config (flocking.rs:15-34); Duration as secs + nanos.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct fp_config {
    pub dt: f32,
    pub avoidance_factor: f32,
    pub centering_factor: f32,
    pub velocity_matching_factor: f32,
    pub distance_weight_threshold: f32,
    pub distance_weight_threshold_falloff: f32,
    pub max_sight_angle: f32,
    pub max_sight_angle_to_lead_boid: f32,
    pub time_to_start_steering_secs: u64,
    pub time_to_start_steering_nanos: u32,
    pub steering_overrides: i32,
}

#[repr(C)]
pub struct fp_flock {
    _private: [u8; 0],
}
#[repr(C)]
pub struct fp_state {
    _private: [u8; 0],
}

extern "C" {
    pub fn fp_config_default(cfg: *mut fp_config) -> c_int;
    pub fn fp_flock_create(out: *mut *mut fp_flock, cfg: *const fp_config, n: u64, state_aos6: *const f32,
                           device: c_int) -> c_int;
    pub fn fp_flock_destroy(f: *mut fp_flock) -> c_int;
    pub fn fp_flock_set_config(f: *mut fp_flock, cfg: *const fp_config) -> c_int;
    pub fn fp_flock_get_config(f: *mut fp_flock, cfg: *mut fp_config) -> c_int;
    pub fn fp_flock_set_method(f: *mut fp_flock, method: c_int) -> c_int;
    pub fn fp_flock_get_method(f: *mut fp_flock, method_in_use: *mut c_int) -> c_int;
    pub fn fp_flock_set_numerics(f: *mut fp_flock, numerics: c_int) -> c_int;
    pub fn fp_flock_get_numerics(f: *mut fp_flock, numerics: *mut c_int, in_use: *mut c_int) -> c_int;
    pub fn fp_flock_set_leads(f: *mut fp_flock, n_leads: u32, leads7: *const f32) -> c_int;
    pub fn fp_flock_set_attractors(f: *mut fp_flock, n: u32, attractors4: *const f32) -> c_int;
    pub fn fp_flock_set_obstacles(f: *mut fp_flock, n: u32, obstacles4: *const f32) -> c_int;
    pub fn fp_flock_set_bbox(f: *mut fp_flock, bbox6: *const f32) -> c_int;
    pub fn fp_flock_set_lead_table(f: *mut fp_flock, steps: u32, n_leads: u32, table7: *const f32) -> c_int;
    pub fn fp_flock_step(f: *mut fp_flock, nsteps: u32) -> c_int;
    pub fn fp_flock_sync(f: *mut fp_flock) -> c_int;
    pub fn fp_flock_read_state(f: *mut fp_flock, out_aos6: *mut f32) -> c_int;
    pub fn fp_flock_write_state(f: *mut fp_flock, state_aos6: *const f32) -> c_int;
    pub fn fp_flock_len(f: *const fp_flock) -> u64;
    pub fn fp_flock_status(f: *mut fp_flock, flags: *mut u32) -> c_int;
    pub fn fp_flock_read_instances(f: *mut fp_flock, out8: *mut f32) -> c_int;
    pub fn fp_flock_read_instances_raw(f: *mut fp_flock, out25: *mut f32) -> c_int;
    pub fn fp_flock_export_instances(f: *mut fp_flock, dst_device_visible: *mut c_void, raw: c_int) -> c_int;
    pub fn fp_flock_read_accel(f: *mut fp_flock, out_accel3: *mut f32, out_comp15: *mut f32) -> c_int;
    pub fn fp_flock_read_neighbors(f: *mut fp_flock, out_count: *mut u32, out_hash: *mut u64) -> c_int;
    pub fn fp_flock_pair_census(f: *mut fp_flock, out4: *mut u64) -> c_int;
    pub fn fp_flock_set_grid_domain(f: *mut fp_flock, lo3: *const f32, hi3: *const f32) -> c_int;
    pub fn fp_flock_grid_info(f: *mut fp_flock, dims3: *mut u32, cell_size: *mut f32, key_bits: *mut u32) -> c_int;
    pub fn fp_flock_set_rebin(f: *mut fp_flock, skin: f32, plan_scale: f32) -> c_int;
    pub fn fp_flock_rebin_info(f: *mut fp_flock, skin: *mut f32, grid_steps: *mut u64, rebins: *mut u64,
                               replayed: *mut u64) -> c_int;
    pub fn fp_flock_shard_info(f: *mut fp_flock, rank: *mut c_int, world: *mut c_int,
                               peer_mapped: *mut c_int) -> c_int;
    pub fn fp_flock_device_state(f: *mut fp_flock, pos4: *mut *const c_void, vel4: *mut *const c_void) -> c_int;
    pub fn fp_flock_timing_begin(f: *mut fp_flock) -> c_int;
    pub fn fp_flock_timing_end(f: *mut fp_flock, steps: *mut u32, span_ms: *mut f32, sort_ms: *mut f32,
                               influence_ms: *mut f32) -> c_int;
    pub fn fp_flock_state_euler(f: *mut fp_flock, h: f32) -> c_int;
    pub fn fp_flock_state_rk4(f: *mut fp_flock, h: f32) -> c_int;
    pub fn fp_state_euler_combine(device: c_int, n: usize, s: *const f32, ds: *const f32, h: f32,
                                  out: *mut f32) -> c_int;
    pub fn fp_state_rk4_combine(device: c_int, n: usize, s: *const f32, k1: *const f32, k2: *const f32,
                                k3: *const f32, k4: *const f32, h: f32, out: *mut f32) -> c_int;
    pub fn fp_state_num_state_elements(kind: c_int) -> c_int;
    pub fn fp_state_create(out: *mut *mut fp_state, device: c_int, kind: c_int, n_elements: u64,
                           state: *const f32) -> c_int;
    pub fn fp_state_destroy(s: *mut fp_state) -> c_int;
    pub fn fp_state_len(s: *const fp_state) -> u64;
    pub fn fp_state_write(s: *mut fp_state, state: *const f32) -> c_int;
    pub fn fp_state_read(s: *mut fp_state, out: *mut f32) -> c_int;
    pub fn fp_state_derivative(s: *mut fp_state, out: *mut f32) -> c_int;
    pub fn fp_state_euler_step(s: *mut fp_state, h: f32, nsteps: u32) -> c_int;
    pub fn fp_state_rk4_step(s: *mut fp_state, h: f32, nsteps: u32) -> c_int;
    pub fn fp_state_sync(s: *mut fp_state) -> c_int;
    pub fn fp_state_device_vector(s: *mut fp_state, dev: *mut *const f32) -> c_int;
    pub fn fp_state_time_steps(s: *mut fp_state, h: f32, rk4: c_int, launches: u32, ms_total: *mut f32) -> c_int;
    pub fn fp_sph_neighbors(device: c_int, n: u64, pos3: *const f32, k: u32, kernal_max_distance: f32,
                            particle_mass: f32, out_index: *mut u32, out_count: *mut u32, out_density: *mut f32,
                            kernel_ms: *mut f32) -> c_int;
    pub fn fp_nccl_unique_id(out128: *mut u8) -> c_int;
    pub fn fp_flock_create_sharded(out: *mut *mut fp_flock, cfg: *const fp_config, n_global: u64,
                                   first_index: u64, n_local: u64, state_aos6: *const f32, device: c_int,
                                   rank: c_int, world: c_int, nccl_unique_id: *const u8) -> c_int;
    pub fn fp_flock_local_len(f: *mut fp_flock, n_local: *mut u64) -> c_int;
    pub fn fp_flock_read_local(f: *mut fp_flock, out_index: *mut u64, out_aos6: *mut f32) -> c_int;
    pub fn fp_flock_write_local(f: *mut fp_flock, n_local: u64, index: *const u64, state_aos6: *const f32) -> c_int;
    pub fn fp_debug_fastmath_check(device: c_int, n: u64, seed: u64, out_mismatch: *mut u64) -> c_int;
    pub fn fp_last_error() -> *const c_char;
    pub fn fp_version() -> *const c_char;
    pub fn fp_launch_count() -> u64;
}
