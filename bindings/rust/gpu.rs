//! GPU Metropolis integrator for vegas (B200, `libvegas_gpu.so`).  SOURCE ONLY: never compiled (no Rust toolchain in
//! the build image) -- see bindings/rust/README.md.  Entry points: include/vegas_gpu.h, include/vegas_host.h.
#![allow(non_camel_case_types)]

use crate::{
    energy::Hamiltonian,
    integrator::Integrator,
    state::{HeisenbergSpin, IsingSpin, Spin, State},
    thermostat::Thermostat,
};
use rand::Rng;
use std::ffi::{CStr, c_char, c_int, c_void};
use std::rc::Rc;

// ---------------------------------------------------------------------------------------- include/vegas_gpu.h
#[repr(C)]
pub struct vegas_model_desc {
    pub model: c_int,      // 0 Ising, 1 Heisenberg
    pub proposal: c_int,   // 0 flip (MetropolisFlipIntegrator), 1 random (MetropolisIntegrator)
    pub precision: c_int,  // 0 f32, 1 f64 (Heisenberg storage)
    pub exchange: f64,
    pub has_exchange: c_int,
    pub has_zeeman: c_int,
    pub has_anisotropy: c_int,
    pub anisotropy_k: f64,
    pub anisotropy_axis: [f64; 3],
    pub has_gauge: c_int,
    pub gauge: f64,
    pub seed: u64,
    pub device: c_int,
    pub force_general: c_int,
}

#[repr(C)]
pub struct vegas_lattice_desc {
    pub unitcell: c_int, // 0 sc, 1 bcc, 2 fcc
    pub nx: u64,
    pub ny: u64,
    pub nz: u64,
    pub pbc_x: c_int,
    pub pbc_y: c_int,
    pub pbc_z: c_int,
    pub literal_from_lattice_filter: c_int,
    pub nz_global: u64,
    pub z_offset: u64,
}

type vegas_gpu_t = *mut c_void;
type vegas_machine_t = *mut c_void;
type stat_cb = extern "C" fn(*mut c_void, *const c_char, f64, f64, f64, f64, f64, f64, f64);
type reduce_cb = extern "C" fn(*mut c_void, *mut f64, u64) -> c_int;   // in-place sum over the ranks of a slab group

unsafe extern "C" {
    fn vegas_gpu_create_lattice(md: *const vegas_model_desc, ld: *const vegas_lattice_desc, out: *mut vegas_gpu_t) -> c_int;
    fn vegas_gpu_destroy(h: vegas_gpu_t);
    fn vegas_gpu_last_error(h: vegas_gpu_t) -> *const c_char;
    fn vegas_gpu_set_thermostat(h: vegas_gpu_t, t: f64, dir: *const f64, mag: f64) -> c_int;
    fn vegas_gpu_step_host_ising(h: vegas_gpu_t, s: *mut i8, n: u64, e: *mut f64, m: *mut f64) -> c_int;
    fn vegas_gpu_step_host_heisenberg(h: vegas_gpu_t, s: *mut f64, n: u64, e: *mut f64, m: *mut f64) -> c_int;
    fn vegas_gpu_upload_ising(h: vegas_gpu_t, s: *const i8, n: u64) -> c_int;
    fn vegas_gpu_upload_heisenberg(h: vegas_gpu_t, s: *const f64, n: u64) -> c_int;
    fn vegas_gpu_total_energy(h: vegas_gpu_t, out: *mut f64) -> c_int;
    fn vegas_gpu_site_energies(h: vegas_gpu_t, out: *mut f64) -> c_int;
    fn vegas_gpu_n_sites(h: vegas_gpu_t) -> u64;
    // include/vegas_host.h
    fn vegas_machine_create(g: vegas_gpu_t, out: *mut vegas_machine_t) -> c_int;
    fn vegas_machine_destroy(m: vegas_machine_t);
    fn vegas_machine_last_error(m: vegas_machine_t) -> *const c_char;
    fn vegas_machine_add_stat_sensor(m: vegas_machine_t, cb: stat_cb, user: *mut c_void) -> c_int;
    fn vegas_machine_set_group(m: vegas_machine_t, reduce: reduce_cb, user: *mut c_void, n_sites_global: u64) -> c_int;
    fn vegas_program_relax(m: vegas_machine_t, steps: u64, temperature: f64) -> c_int;
    fn vegas_program_cooldown(m: vegas_machine_t, tmax: f64, tmin: f64, rate: f64, relax: u64, steps: u64) -> c_int;
    fn vegas_program_hysteresis(m: vegas_machine_t, steps: u64, relax: u64, t: f64, max_field: f64, step: f64) -> c_int;
}

/// Non-zero status of the C ABI with the handle's message (maps onto a new `MachineError::Gpu(String)`).
#[derive(Debug)]
pub struct GpuError(pub i32, pub String);

struct Handle(vegas_gpu_t);
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { vegas_gpu_destroy(self.0) }
    }
}

fn check(rc: c_int, h: vegas_gpu_t) -> Result<(), GpuError> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { vegas_gpu_last_error(h) };
    let text = if msg.is_null() { String::new() } else { unsafe { CStr::from_ptr(msg) }.to_string_lossy().into_owned() };
    Err(GpuError(rc, text))
}

/// Integrator + Hamiltonian in one ref-counted device handle (`Hamiltonian: Clone`, src/energy.rs:45).
#[derive(Clone)]
pub struct GpuMetropolis {
    h: Rc<Handle>,
}

pub enum UnitCell {
    Sc = 0,
    Bcc = 1,
    Fcc = 2,
}

impl GpuMetropolis {
    /// `hamiltonian!(Exchange::from_lattice(exchange, &lattice), Zeeman::new())` (src/input.rs:271) on
    /// `Lattice::{sc,bcc,fcc}(..).expand(x, y, z)` with the given periodicity (src/input.rs:296-322).
    pub fn from_lattice(heisenberg: bool, cell: UnitCell, size: (u64, u64, u64), pbc: (bool, bool, bool), exchange: f64,
                        seed: u64) -> Result<Self, GpuError> {
        let md = vegas_model_desc {
            model: heisenberg as c_int,
            proposal: heisenberg as c_int, // src/input.rs:347-367: Ising -> flip, Heisenberg -> random proposal
            precision: 0,
            exchange,
            has_exchange: 1,
            has_zeeman: 1,
            has_anisotropy: 0,
            anisotropy_k: 0.0,
            anisotropy_axis: [0.0, 0.0, 1.0],
            has_gauge: 0,
            gauge: 0.0,
            seed,
            device: 0,
            force_general: 0,
        };
        let ld = vegas_lattice_desc {
            unitcell: cell as c_int,
            nx: size.0,
            ny: size.1,
            nz: size.2,
            pbc_x: pbc.0 as c_int,
            pbc_y: pbc.1 as c_int,
            pbc_z: pbc.2 as c_int,
            literal_from_lattice_filter: 0,
            nz_global: 0,
            z_offset: 0,
        };
        let mut h: vegas_gpu_t = std::ptr::null_mut();
        check(unsafe { vegas_gpu_create_lattice(&md, &ld, &mut h) }, std::ptr::null_mut())?;
        Ok(Self { h: Rc::new(Handle(h)) })
    }

    fn set_thermostat<S: Spin>(&self, th: &Thermostat<S>) {
        let o = th.field().orientation();
        let dir = [o.sx(), o.sy(), o.sz()];
        // Field::magnitude() is already |magnitude| (src/state.rs:219-221)
        unsafe { vegas_gpu_set_thermostat(self.h.0, th.temperature(), dir.as_ptr(), th.field().magnitude()) };
    }

    /// Device-resident programs (include/vegas_host.h): the State stays in HBM, `on_line` receives the StatSensor
    /// lines (src/instrument.rs:98-131).  `stages` = (name, parameters) as parsed from the TOML input.
    pub fn cooldown(&self, tmax: f64, tmin: f64, rate: f64, relax: u64, steps: u64, on_line: &mut dyn FnMut(&str)) -> Result<(), GpuError> {
        extern "C" fn tramp(user: *mut c_void, line: *const c_char, _t: f64, _f: f64, _e: f64, _cv: f64, _m: f64, _chi: f64, _u4: f64) {
            let f = unsafe { &mut *(user as *mut &mut dyn FnMut(&str)) };
            f(&unsafe { CStr::from_ptr(line) }.to_string_lossy());
        }
        let mut m: vegas_machine_t = std::ptr::null_mut();
        check(unsafe { vegas_machine_create(self.h.0, &mut m) }, self.h.0)?;
        let mut cb: &mut dyn FnMut(&str) = on_line;
        let rc = unsafe {
            vegas_machine_add_stat_sensor(m, tramp, &mut cb as *mut _ as *mut c_void);
            vegas_program_cooldown(m, tmax, tmin, rate, relax, steps)
        };
        let res = if rc == 0 {
            Ok(())
        } else {
            let msg = unsafe { CStr::from_ptr(vegas_machine_last_error(m)) }.to_string_lossy().into_owned();
            Err(GpuError(rc, msg)) // -10.. -15 = ProgramError::{NoSteps, ZeroTemperature, ..} (src/error.rs:31-46)
        };
        unsafe { vegas_machine_destroy(m) };
        res
    }
}

// ------------------------------------------------------------------- Integrator::step, src/integrator.rs:40-49
impl Integrator<IsingSpin> for GpuMetropolis {
    fn step<R: Rng, H: Hamiltonian<IsingSpin>>(&self, _rng: &mut R, thermostat: &Thermostat<IsingSpin>, _hamiltonian: &H,
                                               state: State<IsingSpin>) -> State<IsingSpin> {
        self.set_thermostat(thermostat);
        let mut raw: Vec<i8> = state.spins().iter().map(|s| if *s == IsingSpin::Up { 1 } else { -1 }).collect();
        let rc = unsafe { vegas_gpu_step_host_ising(self.h.0, raw.as_mut_ptr(), raw.len() as u64, std::ptr::null_mut(), std::ptr::null_mut()) };
        assert_eq!(rc, 0, "vegas_gpu_step_host_ising failed"); // the trait has no error channel
        raw.into_iter().map(|s| if s > 0 { IsingSpin::Up } else { IsingSpin::Down }).collect()
    }
}

impl Integrator<HeisenbergSpin> for GpuMetropolis {
    fn step<R: Rng, H: Hamiltonian<HeisenbergSpin>>(&self, _rng: &mut R, thermostat: &Thermostat<HeisenbergSpin>, _hamiltonian: &H,
                                                    state: State<HeisenbergSpin>) -> State<HeisenbergSpin> {
        self.set_thermostat(thermostat);
        let mut raw: Vec<f64> = state.spins().iter().flat_map(|s| [s.sx(), s.sy(), s.sz()]).collect();
        let n = raw.len() / 3;
        let rc = unsafe { vegas_gpu_step_host_heisenberg(self.h.0, raw.as_mut_ptr(), n as u64, std::ptr::null_mut(), std::ptr::null_mut()) };
        assert_eq!(rc, 0, "vegas_gpu_step_host_heisenberg failed");
        raw.chunks_exact(3).map(|c| HeisenbergSpin::from_components(c[0], c[1], c[2])).collect()
    }
}

// ------------------------------------------------------------------------ Hamiltonian, src/energy.rs:45-60
// The instruments call total_energy(thermostat, state) with the host State of the step that just ran; the device holds
// the same state, so the reductions run there (energy convention REFERENCE_COMPOUND = the trait default over the
// compound: exchange counted twice, Zeeman with '+').
impl Hamiltonian<IsingSpin> for GpuMetropolis {
    fn energy(&self, _th: &Thermostat<IsingSpin>, _state: &State<IsingSpin>, index: usize) -> f64 {
        let mut e = vec![0.0f64; unsafe { vegas_gpu_n_sites(self.h.0) } as usize];
        unsafe { vegas_gpu_site_energies(self.h.0, e.as_mut_ptr()) };
        e[index]
    }
    fn total_energy(&self, _th: &Thermostat<IsingSpin>, _state: &State<IsingSpin>) -> f64 {
        let mut e = 0.0;
        unsafe { vegas_gpu_total_energy(self.h.0, &mut e) };
        e
    }
}

impl Hamiltonian<HeisenbergSpin> for GpuMetropolis {
    fn energy(&self, _th: &Thermostat<HeisenbergSpin>, _state: &State<HeisenbergSpin>, index: usize) -> f64 {
        let mut e = vec![0.0f64; unsafe { vegas_gpu_n_sites(self.h.0) } as usize];
        unsafe { vegas_gpu_site_energies(self.h.0, e.as_mut_ptr()) };
        e[index]
    }
    fn total_energy(&self, _th: &Thermostat<HeisenbergSpin>, _state: &State<HeisenbergSpin>) -> f64 {
        let mut e = 0.0;
        unsafe { vegas_gpu_total_energy(self.h.0, &mut e) };
        e
    }
}
