//! `feriphys::simulation::state` (state.rs:1-113) with the integration done on the GPU.
//!
//! Same trait, same struct, same method names and signatures: `Stateful` (four functions,
//! `Vec<f32>` by value), `State<T>::{new, from_state_vector, derivative, as_vector, euler_step,
//! rk4_step, get_elements}`, `Integration`.  A `Stateful` type whose `derivative` the library
//! knows (`device_kind()`, an ADDITION with a default of `None`) is integrated entirely on the
//! device -- one streaming pass per step, every stage of the integrator in registers
//! (fp_state.cu); any other type keeps its own `derivative()` on the host and only the vector
//! arithmetic of state.rs:75-106 runs on the device (`fp_state_*_combine`).
//! NOT COMPILED in this repository (no Rust toolchain in the build image); the tested
//! equivalents are `feriphys_b200/state.py` and `tests/test_gpu_state.py`.
use crate::ffi;
use std::ffi::CStr;

#[derive(Debug, PartialEq, Copy, Clone)]
pub enum Integration {
    Euler,
    Rk4,
}

pub trait Stateful {
    /// Number of f32 elements that are used to represent this object in the State vector.
    fn num_state_elements() -> usize;
    fn from_state_vector(state_data: Vec<f32>) -> Self;
    fn derivative(&self) -> Vec<f32>;
    fn as_state(&self) -> Vec<f32>;
    /// ADDITION: the `ffi::FP_STATEFUL_*` kind whose device kernel evaluates exactly this
    /// `derivative()`; `None` keeps the derivative on the host.
    fn device_kind() -> Option<i32> {
        None
    }
}

fn check(rc: i32) {
    if rc != ffi::FP_OK {
        let msg = unsafe { CStr::from_ptr(ffi::fp_last_error()) }.to_string_lossy().into_owned();
        panic!("feriphys-cuda: {msg}");
    }
}

pub struct State<T: Stateful> {
    elements: Vec<T>,
    device: i32,
}

impl<T: Stateful> State<T> {
    pub fn new(elements: Vec<T>) -> State<T> {
        State { elements, device: 0 }
    }

    /// ADDITION: which CUDA device integrates this state (default 0).
    pub fn on_device(mut self, device: i32) -> State<T> {
        self.device = device;
        self
    }

    pub fn from_state_vector(state_vector: Vec<f32>) -> State<T> {
        let k = T::num_state_elements();
        let elements = state_vector.chunks(k).map(|chunk| T::from_state_vector(chunk.to_vec())).collect();
        State { elements, device: 0 }
    }

    pub fn derivative(&self) -> Vec<f32> {
        self.elements.iter().flat_map(|e| e.derivative()).collect()
    }

    pub fn as_vector(&self) -> Vec<f32> {
        self.elements.iter().flat_map(|e| e.as_state()).collect()
    }

    fn device_step(&self, kind: i32, timestep: f32, rk4: bool) -> State<T> {
        let s = self.as_vector();
        let mut handle = std::ptr::null_mut();
        check(unsafe { ffi::fp_state_create(&mut handle, self.device, kind, self.elements.len() as u64, s.as_ptr()) });
        check(unsafe {
            if rk4 { ffi::fp_state_rk4_step(handle, timestep, 1) } else { ffi::fp_state_euler_step(handle, timestep, 1) }
        });
        let mut out = vec![0f32; s.len()];
        check(unsafe { ffi::fp_state_read(handle, out.as_mut_ptr()) });
        unsafe { ffi::fp_state_destroy(handle) };
        let mut next = State::<T>::from_state_vector(out);
        next.device = self.device;
        next
    }

    /// S_new = S + h * S'  (state.rs:75-83)
    pub fn euler_step(&self, timestep: f32) -> State<T> {
        if let Some(kind) = T::device_kind() {
            return self.device_step(kind, timestep, false);
        }
        let (s, ds) = (self.as_vector(), self.derivative());
        let mut out = vec![0f32; s.len()];
        check(unsafe { ffi::fp_state_euler_combine(self.device, s.len(), s.as_ptr(), ds.as_ptr(), timestep, out.as_mut_ptr()) });
        State::from_state_vector(out).on_device(self.device)
    }

    /// One step of fourth-order Runge-Kutta (state.rs:86-106).
    pub fn rk4_step(&self, timestep: f32) -> State<T> {
        if let Some(kind) = T::device_kind() {
            return self.device_step(kind, timestep, true);
        }
        // the derivative lives on the host: stages as in the reference, the combination on the device
        let s = self.as_vector();
        let stage = |k: &Vec<f32>, h: f32| -> Vec<f32> {
            let mut t = vec![0f32; s.len()];
            check(unsafe { ffi::fp_state_euler_combine(self.device, s.len(), s.as_ptr(), k.as_ptr(), h, t.as_mut_ptr()) });
            State::<T>::from_state_vector(t).derivative()
        };
        let k1 = self.derivative();
        let k2 = stage(&k1, timestep * 0.5);
        let k3 = stage(&k2, timestep * 0.5);
        let k4 = stage(&k3, timestep);
        let mut out = vec![0f32; s.len()];
        check(unsafe {
            ffi::fp_state_rk4_combine(self.device, s.len(), s.as_ptr(), k1.as_ptr(), k2.as_ptr(), k3.as_ptr(),
                                      k4.as_ptr(), timestep, out.as_mut_ptr())
        });
        State::from_state_vector(out).on_device(self.device)
    }

    /// Drops self, returning the State as a Vec<T>.
    pub fn get_elements(self) -> Vec<T> {
        self.elements
    }
}
