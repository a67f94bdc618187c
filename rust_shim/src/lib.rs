//! quokkasim_b200 (Rust host shim) -- `BatchedSimInit` / `BatchedSimulation` for QuokkaSim stock-and-flow models.
//!
//! Stands where `nexosim::SimInit` / `Simulation` stand in the reference's examples
//! (quokkasim_examples/src/bin/discrete_queue.rs:100-111): the same builder calls produce table rows, and the
//! replications are advanced by libquokka_b200.so on the GPU.  There is no CPU execution path.
//!
//! SOURCE ONLY in this repository (no Rust toolchain in the build image); the C ABI it binds is exercised by the
//! Python host side and the test-suite.
pub mod ffi;
pub mod model;

use std::cell::RefCell;
use std::collections::HashMap;
use std::error::Error;
use std::rc::Rc;
use std::time::Duration;

pub use model::*;

/// Absolute time in nanoseconds since the TAI epoch (tai-time's MonotonicTime, reduced to what the examples use).
#[derive(Clone, Copy, Debug, PartialEq, Eq, PartialOrd, Ord)]
pub struct MonotonicTime(pub u64);
impl MonotonicTime {
    pub const EPOCH: MonotonicTime = MonotonicTime(0);
}
impl std::ops::Add<Duration> for MonotonicTime {
    type Output = MonotonicTime;
    fn add(self, d: Duration) -> MonotonicTime {
        MonotonicTime(self.0 + duration_ns(d))
    }
}

type Shared = Rc<RefCell<Component>>;

/// The generated enum of core.rs:504-562 for the discrete family and BasicEnvironment.
#[derive(Clone)]
pub enum ComponentModel {
    StringStock(Shared),
    StringProcess(Shared),
    StringParallelProcess(Shared),
    StringSource(Shared),
    StringSink(Shared),
    BasicEnvironment(Shared),
}

#[allow(non_snake_case)]
impl ComponentModel {
    fn wrap(c: Component) -> Shared {
        Rc::new(RefCell::new(c))
    }
    pub fn StringStock_(c: Component, _mb: Mailbox) -> Self {
        ComponentModel::StringStock(Self::wrap(c))
    }
    pub fn StringProcess_(c: Component, _mb: Mailbox) -> Self {
        ComponentModel::StringProcess(Self::wrap(c))
    }
    pub fn StringParallelProcess_(c: Component, _mb: Mailbox) -> Self {
        ComponentModel::StringParallelProcess(Self::wrap(c))
    }
    pub fn StringSource_(c: Component, _mb: Mailbox) -> Self {
        ComponentModel::StringSource(Self::wrap(c))
    }
    pub fn StringSink_(c: Component, _mb: Mailbox) -> Self {
        ComponentModel::StringSink(Self::wrap(c))
    }
    pub fn BasicEnvironment_(c: Component, _mb: Mailbox) -> Self {
        ComponentModel::BasicEnvironment(Self::wrap(c))
    }
    pub(crate) fn inner(&self) -> &Shared {
        match self {
            ComponentModel::StringStock(c)
            | ComponentModel::StringProcess(c)
            | ComponentModel::StringParallelProcess(c)
            | ComponentModel::StringSource(c)
            | ComponentModel::StringSink(c)
            | ComponentModel::BasicEnvironment(c) => c,
        }
    }
    fn variant(&self) -> &'static str {
        match self {
            ComponentModel::StringStock(_) => "StringStock",
            ComponentModel::StringProcess(_) => "StringProcess",
            ComponentModel::StringParallelProcess(_) => "StringParallelProcess",
            ComponentModel::StringSource(_) => "StringSource",
            ComponentModel::StringSink(_) => "StringSink",
            ComponentModel::BasicEnvironment(_) => "BasicEnvironment",
        }
    }
}
impl std::fmt::Display for ComponentModel {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{}", self.variant())
    }
}

thread_local! {
    /// connect_components! calls in program order: (a, b, n).  Edge order is the subscriber broadcast order.
    static EDGES: RefCell<Vec<(Shared, Shared, i32)>> = RefCell::new(Vec::new());
}

/// connect_components!(&mut a, &mut b[, n]) (core.rs:565-1178).  The pair is validated by the library when the
/// model is lowered (qk_model_connect returns the reference's "No component connection defined from .." text);
/// obviously wrong pairs are rejected here already so the error surfaces at the call site like in the reference.
pub fn connect_components(a: &mut ComponentModel, b: &mut ComponentModel, n: Option<usize>) -> Result<(), Box<dyn Error>> {
    use ComponentModel::*;
    let ok = matches!(
        (&*a, &*b),
        (StringStock(_), StringProcess(_) | StringParallelProcess(_) | StringSink(_))
            | (StringProcess(_) | StringParallelProcess(_) | StringSource(_), StringStock(_))
            | (BasicEnvironment(_), StringProcess(_) | StringParallelProcess(_) | StringSource(_) | StringSink(_))
    );
    if !ok {
        return Err(format!("No component connection defined from {} to {} (n={:?})", a, b, n).into());
    }
    EDGES.with(|e| e.borrow_mut().push((a.inner().clone(), b.inner().clone(), n.map(|x| x as i32).unwrap_or(-1))));
    Ok(())
}
#[macro_export]
macro_rules! connect_components {
    ($a:expr, $b:expr) => {
        $crate::connect_components($a, $b, None)
    };
    ($a:expr, $b:expr, $n:expr) => {
        $crate::connect_components($a, $b, Some($n))
    };
}

/// Loggers only mark components as logged; records are read back per replication.
pub struct Logger {
    pub name: String,
    members: Vec<Shared>,
}
pub enum ComponentLogger {
    StringStockLogger(Logger),
    StringProcessLogger(Logger),
    BasicEnvironmentLogger(Logger),
}
pub struct DiscreteStockLogger;
pub struct DiscreteProcessLogger;
pub struct BasicEnvironmentLogger;
impl DiscreteStockLogger {
    #[allow(clippy::new_ret_no_self)]
    pub fn new(name: &str) -> Logger {
        Logger { name: name.into(), members: Vec::new() }
    }
}
impl DiscreteProcessLogger {
    #[allow(clippy::new_ret_no_self)]
    pub fn new(name: &str) -> Logger {
        Logger { name: name.into(), members: Vec::new() }
    }
}
impl BasicEnvironmentLogger {
    #[allow(clippy::new_ret_no_self)]
    pub fn new(name: &str) -> Logger {
        Logger { name: name.into(), members: Vec::new() }
    }
}
pub fn connect_logger(l: &mut ComponentLogger, c: &mut ComponentModel, n: Option<usize>) -> Result<(), Box<dyn Error>> {
    use ComponentLogger::*;
    use ComponentModel as M;
    let lg = match (l, &*c) {
        (StringStockLogger(lg), M::StringStock(_)) => lg,
        (StringProcessLogger(lg), M::StringProcess(_) | M::StringParallelProcess(_) | M::StringSource(_) | M::StringSink(_)) => lg,
        (BasicEnvironmentLogger(lg), M::BasicEnvironment(_)) => lg,
        _ => return Err(format!("No logger connection defined to {} (n={:?})", c, n).into()),
    };
    c.inner().borrow_mut().logged = true;
    lg.members.push(c.inner().clone());
    Ok(())
}
#[macro_export]
macro_rules! connect_logger {
    ($l:expr, $c:expr) => {
        $crate::connect_logger($l, $c, None)
    };
}

/// register_component!(sim_init, c): registration order = component id = scheduler origin rank - 1.
#[macro_export]
macro_rules! register_component {
    ($sim_init:expr, $c:expr) => {
        $sim_init.add_model(&$c)
    };
}

pub enum ScheduledEvent {
    SetEnvironmentState(BasicEnvironmentState),
}

#[derive(Default)]
pub struct BatchedSimInit {
    components: Vec<Shared>,
    scheduled: Vec<(u64, Shared, u32)>,
}

/// Batch geometry: how many replications, which seeds, which GPU.
#[derive(Clone, Copy, Debug)]
pub struct BatchOptions {
    pub n_replications: u64,
    pub base_seed: u64,
    pub device: i32,
    pub rep_offset: u64,
    pub log_capacity: u32,
}
impl Default for BatchOptions {
    fn default() -> Self {
        BatchOptions { n_replications: 1, base_seed: 0, device: 0, rep_offset: 0, log_capacity: 0 }
    }
}

impl BatchedSimInit {
    pub fn new() -> Self {
        Self::default()
    }
    pub fn add_model(mut self, c: &ComponentModel) -> Self {
        let sh = c.inner().clone();
        sh.borrow_mut().id = Some(self.components.len() as u32);
        self.components.push(sh);
        self
    }
    /// create_scheduled_event!(.., SetEnvironmentState(s), ..) (core.rs:1393-1396)
    pub fn schedule(&mut self, time: MonotonicTime, ev: ScheduledEvent, target: &ComponentModel) -> Result<(), Box<dyn Error>> {
        match (ev, target) {
            (ScheduledEvent::SetEnvironmentState(s), ComponentModel::BasicEnvironment(c)) => {
                self.scheduled.push((time.0, c.clone(), if s == BasicEnvironmentState::Stopped { 1 } else { 0 }));
                Ok(())
            }
            (_, t) => Err(format!("schedule_event not implemented for types (SetEnvironmentState, {})", t).into()),
        }
    }

    /// `sim_init.init(t0)` of the reference, plus the batch geometry.
    pub fn init(self, t0: MonotonicTime, opt: BatchOptions) -> Result<BatchedSimulation, Box<dyn Error>> {
        let mut model: *mut ffi::qk_model = std::ptr::null_mut();
        ffi::check(unsafe { ffi::qk_model_create(&mut model) })?;
        let mut sim = BatchedSimulation {
            model,
            batch: std::ptr::null_mut(),
            dist_ids: HashMap::new(),
            components: self.components.clone(),
            t0: 0,
            initialised: false,
        };
        for sh in &self.components {
            let c = sh.borrow();
            let d = ffi::qk_component_desc {
                kind: c.kind,
                low_capacity: c.low_capacity,
                max_capacity: c.max_capacity,
                n_initial_items: c.resource.len() as u32,
                capacity_hint: 0,
                initial_env_state: if c.state == BasicEnvironmentState::Stopped { 1 } else { 0 },
            };
            let mut id = 0u32;
            ffi::check(unsafe { ffi::qk_model_add_component(model, &d, &mut id) })?;
            debug_assert_eq!(Some(id), c.id);
            let process_like = !matches!(c.kind, ffi::QK_DISCRETE_STOCK | ffi::QK_ENVIRONMENT);
            if process_like {
                for (which, dist) in [(0u32, &c.process_time_distr), (1u32, &c.process_quantity_distr)] {
                    let mut did = 0u32;
                    ffi::check(unsafe { ffi::qk_model_set_dist(model, id, which, &dist.desc, &mut did) })?;
                    sim.dist_ids.entry(dist.uid).or_default().push(did);
                }
                for m in &c.delay_modes {
                    let (mut a, mut b) = (0u32, 0u32);
                    ffi::check(unsafe {
                        ffi::qk_model_add_delay_mode(model, id, &m.until_delay_distr.desc, &m.until_fix_distr.desc, &mut a, &mut b)
                    })?;
                    sim.dist_ids.entry(m.until_delay_distr.uid).or_default().push(a);
                    sim.dist_ids.entry(m.until_fix_distr.uid).or_default().push(b);
                }
            }
            if c.logged {
                ffi::check(unsafe { ffi::qk_model_set_logged(model, id, 1) })?;
            }
        }
        let edges = EDGES.with(|e| e.borrow().clone());
        for (a, b, n) in edges {
            let (ia, ib) = (a.borrow().id, b.borrow().id);
            if let (Some(ia), Some(ib)) = (ia, ib) {
                ffi::check(unsafe { ffi::qk_model_connect(model, ia, ib, n) })?;
            }
        }
        for (t, env, state) in &self.scheduled {
            ffi::check(unsafe { ffi::qk_model_schedule_env_state(model, *t, env.borrow().id.unwrap(), *state) })?;
        }
        let cfg = ffi::qk_batch_config {
            n_reps: opt.n_replications,
            rep_offset: opt.rep_offset,
            base_seed: opt.base_seed,
            device: opt.device,
            log_capacity: opt.log_capacity,
            threads_per_block: 0,
            flags: 0,
            stream: std::ptr::null_mut(),
            lanes_per_replication: 0,
            reserved: 0,
        };
        ffi::check(unsafe { ffi::qk_batch_create(model, &cfg, &mut sim.batch) })?;
        sim.t0 = t0.0; // Model::init runs with the first step, so that tapes set in between feed the init-time draws
        Ok(sim)
    }
}

/// Many independent replications of one model, advanced in lockstep on one GPU (`&mut Simulation` semantics:
/// single owner, not thread-safe).
pub struct BatchedSimulation {
    model: *mut ffi::qk_model,
    batch: *mut ffi::qk_batch,
    dist_ids: HashMap<u64, Vec<u32>>,
    components: Vec<Shared>,
    t0: u64,
    initialised: bool,
}

impl BatchedSimulation {
    fn ensure_init(&mut self) -> Result<(), Box<dyn Error>> {
        if !self.initialised {
            ffi::check(unsafe { ffi::qk_batch_init(self.batch, self.t0) })?;
            self.initialised = true;
        }
        Ok(())
    }
    pub fn step_until(&mut self, t: MonotonicTime) -> Result<(), Box<dyn Error>> {
        self.ensure_init()?;
        ffi::check(unsafe { ffi::qk_batch_step_until(self.batch, t.0) })
    }
    pub fn step_events(&mut self, n: u64) -> Result<(), Box<dyn Error>> {
        self.ensure_init()?;
        ffi::check(unsafe { ffi::qk_batch_step_events(self.batch, n) })
    }
    /// Pre-drawn variate tape: `values[rep * per_rep + i]` is what the i-th `.sample()` of `dist` returns in
    /// replication `rep`.  Must be called before the first step (Model::init runs with the first step).
    pub fn set_tape(&mut self, dist: &Distribution, values: &[f64], per_rep: u64) -> Result<(), Box<dyn Error>> {
        for did in self.dist_ids.get(&dist.uid).cloned().unwrap_or_default() {
            ffi::check(unsafe { ffi::qk_batch_set_tape(self.batch, did, values.as_ptr(), per_rep) })?;
        }
        Ok(())
    }
    /// (status, now_ns, n_events) of one replication; status != 0 is a per-replication fault such as the
    /// reference's panic "Time until next event is zero!" (discrete.rs:541).
    pub fn status(&mut self, rep: u64) -> Result<(u32, u64, u64), Box<dyn Error>> {
        self.ensure_init()?;
        let (mut st, mut now, mut ev) = (0u32, 0u64, 0u64);
        ffi::check(unsafe { ffi::qk_batch_read_status(self.batch, rep, 1, &mut st, &mut now, &mut ev) })?;
        Ok((st, now, ev))
    }
    pub fn comp_stats(&mut self, c: &ComponentModel, rep: u64) -> Result<ffi::qk_comp_stats, Box<dyn Error>> {
        let mut out = ffi::qk_comp_stats::default();
        let id = c.inner().borrow().id.ok_or("component was never registered")?;
        ffi::check(unsafe { ffi::qk_batch_comp_stats(self.batch, rep, 1, id, &mut out) })?;
        Ok(out)
    }
    /// All retained log records of one replication in canonical order (needs `log_capacity > 0`).
    pub fn read_log(&mut self, rep: u64) -> Result<Vec<ffi::qk_log_record>, Box<dyn Error>> {
        let (mut n, mut total) = (0u64, 0u64);
        ffi::check(unsafe { ffi::qk_batch_read_log(self.batch, rep, std::ptr::null_mut(), 0, &mut n, &mut total) })?;
        let mut buf = vec![ffi::qk_log_record::default(); n.max(1) as usize];
        ffi::check(unsafe { ffi::qk_batch_read_log(self.batch, rep, buf.as_mut_ptr(), buf.len() as u64, &mut n, &mut total) })?;
        buf.truncate(n as usize);
        Ok(buf)
    }
    /// EventId "{code}_{seq:06}" of a record (discrete.rs:237)
    pub fn event_id(&self, comp: u16, seq: u32) -> String {
        match comp {
            0xFFFF => "INIT_000000".to_string(),
            0xFFFE => "SCH_000000".to_string(),
            c => format!("{}_{:06}", self.components[c as usize].borrow().element_code, seq),
        }
    }
}

impl Drop for BatchedSimulation {
    fn drop(&mut self) {
        unsafe {
            if !self.batch.is_null() {
                ffi::qk_batch_destroy(self.batch);
            }
            if !self.model.is_null() {
                ffi::qk_model_destroy(self.model);
            }
        }
    }
}
