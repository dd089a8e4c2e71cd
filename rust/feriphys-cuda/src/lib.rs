//! GPU drop-in for `feriphys::simulation::flocking::Simulation` (flocking.rs:53-246).
//!
//! Same public surface -- `new`, `step`, `get_timestep`, `sync_sim_config_from_ui`,
//! `get_boid_instances` -- over `libferiphys_cuda.so`.  NOT COMPILED in this repository
//! (no Rust toolchain in the build image); kept mechanical on purpose.  The tested
//! equivalents of the host logic below are `cpp/feriphys_cuda.hpp` and
//! `feriphys_b200/flocking.py`.
pub mod ffi;
pub mod state;

use cgmath::{Quaternion, Vector3, Zero};
use state::Stateful;
use std::{ffi::CStr, ptr, time::Duration};

/// flocking::Config (flocking.rs:15-51), field for field.
pub struct Config {
    pub dt: f32,
    pub avoidance_factor: f32,
    pub centering_factor: f32,
    pub velocity_matching_factor: f32,
    pub distance_weight_threshold: f32,
    pub distance_weight_threshold_falloff: f32,
    pub max_sight_angle: f32,
    pub max_sight_angle_to_lead_boid: f32,
    pub time_to_start_steering: Duration,
    pub steering_overrides: bool,
}

impl Default for Config {
    fn default() -> Self {
        Self {
            dt: Duration::from_millis(1).as_secs_f32(),
            avoidance_factor: 1.0,
            centering_factor: 0.1,
            velocity_matching_factor: 0.5,
            distance_weight_threshold: 15.0,
            distance_weight_threshold_falloff: 1.0,
            max_sight_angle: std::f32::consts::PI / 2.0,
            max_sight_angle_to_lead_boid: std::f32::consts::PI,
            time_to_start_steering: Duration::from_secs(4),
            steering_overrides: false,
        }
    }
}

impl Config {
    fn to_c(&self) -> ffi::fp_config {
        ffi::fp_config {
            dt: self.dt,
            avoidance_factor: self.avoidance_factor,
            centering_factor: self.centering_factor,
            velocity_matching_factor: self.velocity_matching_factor,
            distance_weight_threshold: self.distance_weight_threshold,
            distance_weight_threshold_falloff: self.distance_weight_threshold_falloff,
            max_sight_angle: self.max_sight_angle,
            max_sight_angle_to_lead_boid: self.max_sight_angle_to_lead_boid,
            time_to_start_steering_secs: self.time_to_start_steering.as_secs(),
            time_to_start_steering_nanos: self.time_to_start_steering.subsec_nanos(),
            steering_overrides: self.steering_overrides as i32,
        }
    }
}

/// parametric.rs:4-22
pub struct Parametric {
    path: fn(t: f32) -> Vector3<f32>,
    curr_time: f32,
}
impl Parametric {
    pub fn new(path: fn(t: f32) -> Vector3<f32>) -> Parametric {
        Parametric { path, curr_time: 0.0 }
    }
    pub fn step(&mut self, dt: f32) -> Vector3<f32> {
        let position = (self.path)(self.curr_time);
        self.curr_time = self.curr_time + dt;
        position
    }
}

/// boid.rs:13-53 -- stays on the host: a Rust fn pointer cannot cross the FFI.
pub struct LeadBoid {
    parametric: Parametric,
    position: Vector3<f32>,
    velocity: Vector3<f32>,
    weight: f32,
}
impl LeadBoid {
    pub fn new(path: fn(t: f32) -> Vector3<f32>) -> LeadBoid {
        LeadBoid { parametric: Parametric::new(path), position: path(0.0), velocity: Vector3::zero(), weight: 10.0 }
    }
    pub fn step(&mut self, dt: Duration) {
        if dt.is_zero() {
            return;
        }
        let new_position = self.parametric.step(dt.as_secs_f32());
        self.velocity = (new_position - self.position) / dt.as_secs_f32();
        self.position = new_position;
    }
    fn row(&self) -> [f32; 7] {
        [self.position.x, self.position.y, self.position.z, self.velocity.x, self.velocity.y, self.velocity.z,
         self.weight]
    }
}

#[derive(Clone)]
pub struct Obstacle {
    pub position: Vector3<f32>,
    pub radius: f32,
}
pub struct PointAttractor {
    pub position: Vector3<f32>,
    pub mass: f32,
}
pub struct BoundingBox {
    pub x_range: std::ops::Range<f32>,
    pub y_range: std::ops::Range<f32>,
    pub z_range: std::ops::Range<f32>,
}
/// graphics/instance.rs:7-11
pub struct Instance {
    pub position: Vector3<f32>,
    pub rotation: Quaternion<f32>,
    pub scale: f32,
}

/// boid.rs:56-64 with the acceleration of the step carried as frozen state -- the pattern of the
/// reference's own Stateful types (springy_mesh.rs:199-257): `[p, v, a]`, derivative `[v, a, 0]`.
#[derive(Clone, Copy, Debug, PartialEq)]
pub struct FlockingBoid {
    pub position: Vector3<f32>,
    pub velocity: Vector3<f32>,
    pub acceleration: Vector3<f32>,
}
impl Stateful for FlockingBoid {
    fn num_state_elements() -> usize {
        9
    }
    fn from_state_vector(s: Vec<f32>) -> Self {
        if s.len() != Self::num_state_elements() {
            panic!("State Vector incorrect size!")
        }
        FlockingBoid {
            position: Vector3::new(s[0], s[1], s[2]),
            velocity: Vector3::new(s[3], s[4], s[5]),
            acceleration: Vector3::new(s[6], s[7], s[8]),
        }
    }
    fn derivative(&self) -> Vec<f32> {
        vec![self.velocity.x, self.velocity.y, self.velocity.z, self.acceleration.x, self.acceleration.y,
             self.acceleration.z, 0.0, 0.0, 0.0]
    }
    fn as_state(&self) -> Vec<f32> {
        vec![self.position.x, self.position.y, self.position.z, self.velocity.x, self.velocity.y, self.velocity.z,
             self.acceleration.x, self.acceleration.y, self.acceleration.z]
    }
    fn device_kind() -> Option<i32> {
        Some(ffi::FP_STATEFUL_BOID)
    }
}

fn check(rc: i32) {
    if rc != ffi::FP_OK {
        // the reference's failure mode is panic!/unwrap
        let msg = unsafe { CStr::from_ptr(ffi::fp_last_error()) }.to_string_lossy().into_owned();
        panic!("feriphys-cuda: {msg}");
    }
}

pub struct Simulation {
    config: Config,
    handle: *mut ffi::fp_flock,
    n: usize,
    lead_boids: Option<Vec<LeadBoid>>,
}

impl Simulation {
    /// Simulation::new (flocking.rs:63-95): jitter drawn here with `rand::random`, exactly as
    /// the reference does, then handed to the library as explicit state.
    pub fn new(
        initial_positions: Vec<Vector3<f32>>,
        num_boids: u32,
        bounding_box: Option<BoundingBox>,
        lead_boids: Option<Vec<LeadBoid>>,
        obstacles: Option<Vec<Obstacle>>,
        attractors: Option<Vec<PointAttractor>>,
        jitter: &mut dyn FnMut() -> f32, // `rand::random::<f32>` in the demo
    ) -> Simulation {
        let mut state = Vec::<f32>::with_capacity(num_boids as usize * 6);
        for position in &initial_positions {
            for _ in 0..num_boids / initial_positions.len() as u32 {
                state.extend([position.x + jitter(), position.y + jitter(), position.z + jitter()]);
                state.extend([jitter(), jitter(), jitter()]);
            }
        }
        Self::from_state(state, bounding_box, lead_boids, obstacles, attractors, 0)
    }

    /// ADDITION: explicit state, `[px py pz vx vy vz]` per boid.
    pub fn from_state(
        state: Vec<f32>,
        bounding_box: Option<BoundingBox>,
        lead_boids: Option<Vec<LeadBoid>>,
        obstacles: Option<Vec<Obstacle>>,
        attractors: Option<Vec<PointAttractor>>,
        device: i32, // CUDA ordinal
    ) -> Simulation {
        let config = Config::default();
        let n = state.len() / 6;
        let mut handle = ptr::null_mut();
        let c = config.to_c();
        check(unsafe { ffi::fp_flock_create(&mut handle, &c, n as u64, state.as_ptr(), device) });
        if let Some(b) = &bounding_box {
            let r = [b.x_range.start, b.x_range.end, b.y_range.start, b.y_range.end, b.z_range.start, b.z_range.end];
            check(unsafe { ffi::fp_flock_set_bbox(handle, r.as_ptr()) });
        }
        if let Some(o) = &obstacles {
            let t: Vec<f32> = o.iter().flat_map(|x| [x.position.x, x.position.y, x.position.z, x.radius]).collect();
            check(unsafe { ffi::fp_flock_set_obstacles(handle, o.len() as u32, t.as_ptr()) });
        }
        if let Some(a) = &attractors {
            let t: Vec<f32> = a.iter().flat_map(|x| [x.position.x, x.position.y, x.position.z, x.mass]).collect();
            check(unsafe { ffi::fp_flock_set_attractors(handle, a.len() as u32, t.as_ptr()) });
        }
        let sim = Simulation { config, handle, n, lead_boids };
        sim.push_leads();
        sim
    }

    fn push_leads(&self) {
        let rows: Vec<f32> = self.lead_boids.iter().flatten().flat_map(|l| l.row()).collect();
        let n = self.lead_boids.as_ref().map_or(0, |l| l.len()) as u32;
        check(unsafe { ffi::fp_flock_set_leads(self.handle, n, if n == 0 { ptr::null() } else { rows.as_ptr() }) });
    }

    /// flocking.rs:97-131
    pub fn step(&mut self) -> Duration {
        if self.lead_boids.is_some() {
            self.push_leads();
        }
        check(unsafe { ffi::fp_flock_step(self.handle, 1) });
        if let Some(lead_boids) = &mut self.lead_boids {
            for lead_boid in lead_boids.iter_mut() {
                lead_boid.step(Duration::from_secs_f32(self.config.dt));
            }
        }
        self.get_timestep()
    }

    pub fn get_timestep(&self) -> Duration {
        Duration::from_secs_f32(self.config.dt)
    }

    /// flocking.rs:215-228 -- `ui_config` is what `FlockingUi::get_gui_state_mut()` returns.
    pub fn sync_sim_config_from_ui(&mut self, ui_config: Config) {
        self.config = ui_config;
        let c = self.config.to_c();
        check(unsafe { ffi::fp_flock_set_config(self.handle, &c) });
    }

    /// flocking.rs:230-245
    pub fn get_boid_instances(&self) -> Vec<Instance> {
        let mut raw = vec![0f32; self.n * 8];
        check(unsafe { ffi::fp_flock_read_instances(self.handle, raw.as_mut_ptr()) });
        raw.chunks_exact(8)
            .map(|r| Instance {
                position: Vector3::new(r[0], r[1], r[2]),
                rotation: Quaternion::new(r[3], r[4], r[5], r[6]),
                scale: r[7],
            })
            .collect()
    }

    /// ADDITION: `FP_NUMERICS_EXACT` / `FP_NUMERICS_FAST` (include/feriphys_cuda.h)
    pub fn set_numerics(&mut self, numerics: i32) {
        check(unsafe { ffi::fp_flock_set_numerics(self.handle, numerics) });
    }

    /// ADDITION: the flock as `State<FlockingBoid>` advanced by `State::euler_step(h)` /
    /// `State::rk4_step(h)` (state.rs:75-106) with this step's accelerations frozen, device-resident.
    pub fn state_step(&mut self, integration: state::Integration, h: f32) {
        check(unsafe {
            match integration {
                state::Integration::Euler => ffi::fp_flock_state_euler(self.handle, h),
                state::Integration::Rk4 => ffi::fp_flock_state_rk4(self.handle, h),
            }
        });
    }

    /// ADDITION
    pub fn read_state(&self) -> Vec<f32> {
        let mut s = vec![0f32; self.n * 6];
        check(unsafe { ffi::fp_flock_read_state(self.handle, s.as_mut_ptr()) });
        s
    }
}

impl Drop for Simulation {
    fn drop(&mut self) {
        unsafe { ffi::fp_flock_destroy(self.handle) };
    }
}

// The handle owns device memory only; one thread at a time (`&mut self` on step).
unsafe impl Send for Simulation {}
