"""CPU: the Rust crate under rust/feriphys-cuda cannot be compiled here (no cargo / rustc), so what
CAN drift is checked textually: its `extern "C"` block against include/feriphys_cuda.h (names,
argument counts, every declared entry bound), the source list its build.rs derives from the
Makefile, and the State API the north star asks it to keep (state.rs:4-16, 37-113)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CRATE = os.path.join(ROOT, "rust", "feriphys-cuda")


def _header_decls():
    hdr = open(os.path.join(ROOT, "include", "feriphys_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(fp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def _ffi_decls():
    src = open(os.path.join(CRATE, "src", "ffi.rs")).read()
    out = {}
    for m in re.finditer(r"pub fn (fp_[a-z0-9_]+)\s*\(([^)]*)\)", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return out


def test_ffi_rs_binds_every_header_entry_with_the_same_arity():
    hdr, ffi = _header_decls(), _ffi_decls()
    assert len(hdr) >= 50
    assert set(hdr) == set(ffi), sorted(set(hdr) ^ set(ffi))
    for name, n in hdr.items():
        assert ffi[name] == n, (name, n, ffi[name])


def test_ffi_rs_constants_match_the_header():
    hdr = open(os.path.join(ROOT, "include", "feriphys_cuda.h")).read()
    src = open(os.path.join(CRATE, "src", "ffi.rs")).read()
    consts = dict(re.findall(r"pub const (FP_[A-Z_]+): c_int = (-?\d+);", src))
    assert len(consts) >= 12
    for name, value in consts.items():
        m = re.search(rf"\b{name}\s*=\s*(-?\d+)|#define\s+{name}\s+\(?(-?\d+)\)?", hdr)
        assert m, name
        assert int(m.group(1) or m.group(2)) == int(value), name


def test_build_rs_compiles_exactly_the_makefile_sources():
    mk = open(os.path.join(ROOT, "feriphys_b200", "csrc", "Makefile")).read()
    line = next(l for l in mk.splitlines() if l.strip().startswith("SRCS"))
    srcs = [w for w in line.split("=", 1)[1].split() if w.endswith(".cu")]
    on_disk = sorted(f for f in os.listdir(os.path.join(ROOT, "feriphys_b200", "csrc")) if f.endswith(".cu"))
    assert sorted(srcs) == on_disk, "the Makefile does not list every .cu file of csrc/"
    b = open(os.path.join(CRATE, "build.rs")).read()
    # build.rs parses that same line at build time (no second list to keep in step) ...
    assert 'starts_with("SRCS")' in b and "feriphys_b200/csrc" in b and "Makefile" in b
    assert not re.search(r'"fp_[a-z_]+\.cu"', b), "build.rs must not carry its own list of sources"
    # ... and asks for sm_100a and tracks the headers
    assert "arch=compute_100a,code=sm_100a" in b and "rerun-if-changed" in b and '"cuh"' in b


def test_state_api_keeps_the_reference_surface():
    ref = ["fn num_state_elements() -> usize", "fn from_state_vector(state_data: Vec<f32>) -> Self",
           "fn derivative(&self) -> Vec<f32>", "fn as_state(&self) -> Vec<f32>",
           "pub fn new(elements: Vec<T>) -> State<T>", "pub fn from_state_vector(state_vector: Vec<f32>) -> State<T>",
           "pub fn derivative(&self) -> Vec<f32>", "pub fn as_vector(&self) -> Vec<f32>",
           "pub fn euler_step(&self, timestep: f32) -> State<T>", "pub fn rk4_step(&self, timestep: f32) -> State<T>",
           "pub fn get_elements(self) -> Vec<T>", "pub enum Integration", "pub trait Stateful",
           "pub struct State<T: Stateful>"]
    src = open(os.path.join(CRATE, "src", "state.rs")).read()
    for item in ref:
        assert item in src, item
    lib = open(os.path.join(CRATE, "src", "lib.rs")).read()
    assert "impl Stateful for FlockingBoid" in lib and "pub mod state;" in lib
    assert "device: i32" in lib and "state.as_ptr(), device)" in lib       # the CUDA ordinal is a parameter
