//! Model-builder surface of quokkasim v0.2.2, restated as plain descriptions (no NeXosim ports behind them).
//!
//! Names follow the reference so that example code reads the same:
//!   DistributionConfig / DistributionFactory::create     quokkasim/src/common.rs:37-148
//!   builders of #[derive(WithMethods)]                   quokkasim_derive_macros/src/lib.rs:22-94
//!   DiscreteStock / Process / Source / Sink / ParallelProcess   components/discrete.rs
//!   DelayMode / DelayModeChange                          delays.rs:6-11, 45-50
use std::time::Duration;

use crate::ffi;

/// What `Distribution` has to expose for the batched path: kind, parameters and the seed the factory assigned
/// (the reference's enum owns an opaque `SmallRng`, common.rs:25-32).
#[derive(Clone, Debug)]
pub struct Distribution {
    pub(crate) desc: ffi::qk_dist_desc,
    /// identity of this instance: components that share one `Distribution` value share its variate tape
    pub(crate) uid: u64,
}

fn next_uid() -> u64 {
    use std::sync::atomic::{AtomicU64, Ordering};
    static N: AtomicU64 = AtomicU64::new(1);
    N.fetch_add(1, Ordering::Relaxed)
}

impl Distribution {
    #[allow(non_snake_case)]
    pub fn Constant(value: f64) -> Self {
        Distribution { desc: ffi::qk_dist_desc { kind: ffi::QK_DIST_CONSTANT, stream: 0, p: [value, 0., 0., 0.] }, uid: next_uid() }
    }
}

impl Default for Distribution {
    fn default() -> Self {
        Distribution::Constant(1.) // common.rs:181-185
    }
}

#[derive(Clone, Debug)]
pub enum DistributionConfig {
    Uniform { min: f64, max: f64 },
    Triangular { min: f64, max: f64, mode: f64 },
    Constant(f64),
    Normal { mean: f64, std: f64 },
    TruncNormal { mean: f64, std: f64, min: Option<f64>, max: Option<f64> },
    Exponential { mean: f64 },
}

#[derive(Debug)]
pub struct DistributionParametersError(pub String);
impl std::fmt::Display for DistributionParametersError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        write!(f, "{}", self.0)
    }
}
impl std::error::Error for DistributionParametersError {}

pub struct DistributionFactory {
    pub base_seed: u64,
    pub next_seed: u64,
}

impl DistributionFactory {
    pub fn new(base_seed: u64) -> Self {
        DistributionFactory { base_seed, next_seed: base_seed }
    }

    /// Seeds increment per created distribution, except that the Normal / TruncNormal / Exponential arms of the
    /// reference return early and skip the increment (common.rs:98,121,134 vs :145) -- reproduced, because the
    /// stream id decides which instances share a variate sequence.
    pub fn create(&mut self, config: DistributionConfig) -> Result<Distribution, DistributionParametersError> {
        let stream = self.next_seed as u32;
        let mk = |kind: u32, p: [f64; 4]| Distribution { desc: ffi::qk_dist_desc { kind, stream, p }, uid: next_uid() };
        let (d, bump) = match config {
            DistributionConfig::Uniform { min, max } => (mk(ffi::QK_DIST_UNIFORM, [min, max, 0., 0.]), true),
            DistributionConfig::Triangular { min, max, mode } => {
                if !(max >= mode && mode >= min) || !(max > min) {
                    return Err(DistributionParametersError("invalid Triangular parameters".into()));
                }
                (mk(ffi::QK_DIST_TRIANGULAR, [min, max, mode, 0.]), true)
            }
            DistributionConfig::Constant(v) => (Distribution::Constant(v), true),
            DistributionConfig::Normal { mean, std } => {
                if !(std >= 0.) || !std.is_finite() {
                    return Err(DistributionParametersError("invalid Normal parameters".into()));
                }
                (mk(ffi::QK_DIST_NORMAL, [mean, std, 0., 0.]), false)
            }
            DistributionConfig::TruncNormal { mean, std, min, max } => {
                let (lo, hi) = (min.unwrap_or(f64::MIN), max.unwrap_or(f64::MAX));
                if lo >= hi {
                    return Err(DistributionParametersError(
                        "Minimum value cannot be greater than or equal maximum value".into(),
                    ));
                }
                (mk(ffi::QK_DIST_TRUNCNORMAL, [mean, std, lo, hi]), false)
            }
            DistributionConfig::Exponential { mean } => {
                if !(mean > 0.) {
                    return Err(DistributionParametersError("invalid Exponential parameters".into()));
                }
                (mk(ffi::QK_DIST_EXPONENTIAL, [mean, 0., 0., 0.]), false)
            }
        };
        if bump {
            self.next_seed += 1;
        }
        Ok(d)
    }
}

#[derive(Clone, Debug)]
pub struct DelayMode {
    pub name: String,
    pub until_delay_distr: Distribution,
    pub until_fix_distr: Distribution,
}

pub enum DelayModeChange {
    Add(DelayMode),
    Remove(String),
    RemoveAll,
}

/// discrete.rs:82-103
pub type ItemDeque = std::collections::VecDeque<String>;

/// discrete.rs:664-686: items are "{prefix}_{index:0>num_digits}"
#[derive(Clone, Debug)]
pub struct StringItemFactory {
    pub prefix: String,
    pub next_index: u64,
    pub num_digits: usize,
}
impl Default for StringItemFactory {
    fn default() -> Self {
        StringItemFactory { prefix: "Item".into(), next_index: 0, num_digits: 4 }
    }
}
impl StringItemFactory {
    pub fn item_name(&self, index: u64) -> String {
        format!("{}_{:0>width$}", self.prefix, self.next_index + index, width = self.num_digits)
    }
}

#[derive(Clone, Copy, Debug, PartialEq)]
pub enum BasicEnvironmentState {
    Normal,
    Stopped,
}

/// Every component is one row.  `kind` selects which fields are meaningful.
#[derive(Clone, Debug)]
pub struct Component {
    pub(crate) kind: u32,
    pub element_name: String,
    pub element_code: String,
    pub element_type: String,
    pub low_capacity: u32,
    pub max_capacity: u32,
    pub resource: ItemDeque,
    pub process_time_distr: Distribution,
    pub process_quantity_distr: Distribution,
    pub delay_modes: Vec<DelayMode>,
    pub item_factory: StringItemFactory,
    pub state: BasicEnvironmentState,
    pub(crate) id: Option<u32>,
    pub(crate) logged: bool,
}

impl Component {
    fn of(kind: u32, type_name: &str) -> Self {
        Component {
            kind,
            element_name: type_name.to_string(),
            element_code: String::new(),
            element_type: type_name.to_string(),
            low_capacity: 0,
            max_capacity: 1, // discrete.rs:70-71
            resource: ItemDeque::default(),
            process_time_distr: Distribution::default(),
            process_quantity_distr: Distribution::default(),
            delay_modes: Vec::new(),
            item_factory: StringItemFactory::default(),
            state: BasicEnvironmentState::Normal,
            id: None,
            logged: false,
        }
    }
    pub fn with_name(mut self, name: &str) -> Self {
        self.element_name = name.to_string();
        self
    }
    pub fn with_code(mut self, code: &str) -> Self {
        self.element_code = code.to_string();
        self
    }
    pub fn with_type(mut self, t: &str) -> Self {
        self.element_type = t.to_string();
        self
    }
    pub fn with_low_capacity(mut self, v: u32) -> Self {
        self.low_capacity = v;
        self
    }
    pub fn with_max_capacity(mut self, v: u32) -> Self {
        self.max_capacity = v;
        self
    }
    pub fn with_initial_resource(mut self, r: ItemDeque) -> Self {
        self.resource = r;
        self
    }
    pub fn with_process_time_distr(mut self, d: Distribution) -> Self {
        self.process_time_distr = d;
        self
    }
    pub fn with_process_quantity_distr(mut self, d: Distribution) -> Self {
        self.process_quantity_distr = d;
        self
    }
    pub fn with_item_factory(mut self, f: StringItemFactory) -> Self {
        self.item_factory = f;
        self
    }
    pub fn with_state(mut self, s: BasicEnvironmentState) -> Self {
        self.state = s;
        self
    }
    /// DelayModes::modify (delays.rs:157-174): Add on an existing name keeps its position
    pub fn with_delay_mode(mut self, change: DelayModeChange) -> Self {
        match change {
            DelayModeChange::Add(m) => {
                if let Some(slot) = self.delay_modes.iter_mut().find(|x| x.name == m.name) {
                    *slot = m;
                } else {
                    self.delay_modes.push(m);
                }
            }
            DelayModeChange::Remove(name) => self.delay_modes.retain(|x| x.name != name),
            DelayModeChange::RemoveAll => self.delay_modes.clear(),
        }
        self
    }
}

macro_rules! component_type {
    ($name:ident, $kind:expr) => {
        pub struct $name;
        impl $name {
            #[allow(clippy::new_ret_no_self)]
            pub fn new() -> Component {
                Component::of($kind, stringify!($name))
            }
        }
    };
}
component_type!(DiscreteStock, ffi::QK_DISCRETE_STOCK);
component_type!(DiscreteProcess, ffi::QK_DISCRETE_PROCESS);
component_type!(DiscreteSource, ffi::QK_DISCRETE_SOURCE);
component_type!(DiscreteSink, ffi::QK_DISCRETE_SINK);
component_type!(DiscreteParallelProcess, ffi::QK_DISCRETE_PARALLEL_PROCESS);

pub struct BasicEnvironment;
impl BasicEnvironment {
    #[allow(clippy::new_ret_no_self)]
    pub fn new() -> Component {
        let mut c = Component::of(ffi::QK_ENVIRONMENT, "BasicEnvironment");
        c.element_name = String::new();
        c
    }
}

/// Kept so that `ComponentModel::StringStock(component, Mailbox::new())` reads like the reference; inert here.
pub struct Mailbox;
impl Mailbox {
    #[allow(clippy::new_ret_no_self)]
    pub fn new() -> Mailbox {
        Mailbox
    }
}

pub(crate) fn duration_ns(d: Duration) -> u64 {
    d.as_secs() * 1_000_000_000 + d.subsec_nanos() as u64
}
