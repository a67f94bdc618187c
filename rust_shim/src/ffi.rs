//! `extern "C"` declarations for include/quokka_b200.h (ABI version 2): every entry point of the header, plain old data
//! only.  Never compiled in the image this was written in (no Rust toolchain): kept in step with the header by hand.
#![allow(non_camel_case_types, dead_code)]
use core::ffi::{c_char, c_void};

pub const QK_OK: i32 = 0;

// qk_component_kind
pub const QK_DISCRETE_STOCK: u32 = 1;
pub const QK_DISCRETE_PROCESS: u32 = 2;
pub const QK_DISCRETE_SOURCE: u32 = 3;
pub const QK_DISCRETE_SINK: u32 = 4;
pub const QK_DISCRETE_PARALLEL_PROCESS: u32 = 5;
pub const QK_ENVIRONMENT: u32 = 6;
pub const QK_VECTOR_STOCK: u32 = 16;
pub const QK_VECTOR_PROCESS: u32 = 17;
pub const QK_VECTOR_SOURCE: u32 = 18;
pub const QK_VECTOR_SINK: u32 = 19;
pub const QK_VECTOR_COMBINER: u32 = 20;
pub const QK_VECTOR_SPLITTER: u32 = 21;
// container family (components/vector_container.rs; ComponentModel::Vector3Container*, core.rs:549-555)
pub const QK_CONTAINER_STOCK: u32 = 32;
pub const QK_CONTAINER_PROCESS: u32 = 33;
pub const QK_CONTAINER_SOURCE: u32 = 34;
pub const QK_CONTAINER_SINK: u32 = 35;
pub const QK_CONTAINER_PARALLEL_PROCESS: u32 = 36;
pub const QK_CONTAINER_LOAD_PROCESS: u32 = 37; // max_capacity of the descriptor = max_process_count
pub const QK_CONTAINER_UNLOAD_PROCESS: u32 = 38;
/// first fp64 word of a container's resource when it is `None`
pub const QK_CONTAINER_NONE_BITS: u64 = 0x7FF4_0000_0000_00C0;

// qk_dist_kind
pub const QK_DIST_CONSTANT: u32 = 0;
pub const QK_DIST_UNIFORM: u32 = 1;
pub const QK_DIST_TRIANGULAR: u32 = 2;
pub const QK_DIST_NORMAL: u32 = 3;
pub const QK_DIST_TRUNCNORMAL: u32 = 4;
pub const QK_DIST_EXPONENTIAL: u32 = 5;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct qk_dist_desc {
    pub kind: u32,
    pub stream: u32,
    pub p: [f64; 4],
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct qk_component_desc {
    pub kind: u32,
    pub low_capacity: u32,
    pub max_capacity: u32,
    pub n_initial_items: u32,
    pub capacity_hint: u32,
    pub initial_env_state: u32,
}

/// qk_batch_config.flags (quokka_b200.h): 0 lets the library choose the state placement
pub const QK_BATCH_STATE_IN_HBM: u32 = 1;
pub const QK_BATCH_STATE_ON_CHIP: u32 = 2;
pub const QK_BATCH_STATE_LANE_GROUP: u32 = 4;
pub const QK_BATCH_GROUP_PHASED: u32 = 8;
pub const QK_BATCH_STATE_SPLIT: u32 = 16;
pub const QK_BATCH_QUEUE_TWO_TIER: u32 = 32;

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct qk_batch_config {
    pub n_reps: u64,
    pub rep_offset: u64,
    pub base_seed: u64,
    pub device: i32,
    pub log_capacity: u32,
    pub threads_per_block: u32,
    pub flags: u32,
    pub stream: *mut c_void,
    pub lanes_per_replication: u32, // ABI 2: with QK_BATCH_STATE_LANE_GROUP (4): 4, 8 or 16; 0 = choose
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct qk_log_record {
    pub time_ns: u64,
    pub comp: u16,
    pub kind: u8,
    pub reason: u8,
    pub seq: u32,
    pub src_comp: u16,
    pub aux: u16,
    pub src_seq: u32,
    pub payload: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct qk_comp_stats {
    pub count: u64,
    pub time_ns: u64,
    pub aux: u64,
    pub level: u64,
    pub fvalue: f64,
    pub faux: f64,
}

pub const QK_HIST_BINS: usize = 32;
/// whole-batch (or, after qk_multi_stats, whole-job) reduction of qk_comp_stats over the replications
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct qk_comp_summary {
    pub sum_count: u64,
    pub sum_time_lo: u64,
    pub sum_time_hi: u64,
    pub sum_level: u64,
    pub min_count: u64,
    pub max_count: u64,
    pub min_time: u64,
    pub max_time: u64,
    pub sumsq_time: f64,
    pub sumsq_count: f64,
    pub hist_time: [u64; QK_HIST_BINS],
    pub hist_count: [u64; QK_HIST_BINS],
    pub hist_lo_time: u64,
    pub hist_hi_time: u64,
    pub hist_lo_count: u64,
    pub hist_hi_count: u64,
    pub sumsq_time_ns2: [u64; 3],
    pub sumsq_count_lo: u64,
    pub sumsq_count_hi: u64,
    pub sumsq_level: u64,
}

/// words of the packed partials that cross GPUs (qk_batch_reduce_partials*)
pub const QK_PARTIAL_WORDS: usize = 20;

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct qk_batch_summary {
    pub n_reps: u64,
    pub n_events: u64,
    pub n_faulted: u64,
    pub n_log_records: u64,
    pub min_now_ns: u64,
    pub max_now_ns: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct qk_batch_info {
    pub threads_per_block: u32,
    pub grid_blocks: u32,
    pub shared_bytes_per_block: u32,
    pub table_bytes: u32,
    pub slots64: u32,
    pub slots32: u32,
    pub state_bytes_per_rep: u32,
    pub resident_blocks_per_sm: u32,
    pub state_in_hbm: u32,
    pub cold_bytes_per_rep: u32,
    pub replications_padded: u64,
    pub state_bytes_total: u64,
    pub log_bytes_total: u64,
    pub lanes_per_replication: u32,
    pub wave_blocks: u32,
    pub last_step_chunks: u32,
    pub state_split: u32,
}

/// one job on several GPUs of one box: `devices` null / `n_devices` 0 = every visible device
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct qk_multi_config {
    pub n_reps: u64,
    pub rep_offset: u64,
    pub base_seed: u64,
    pub devices: *const i32,
    pub n_devices: u32,
    pub log_capacity: u32,
    pub threads_per_block: u32,
    pub flags: u32,
    pub lanes_per_replication: u32,
}

/// fields a user-defined resource may have (qk_model_set_vector_n)
pub const QK_VEC_MAX_DIM: usize = 8;

pub enum qk_model {}
pub enum qk_batch {}
pub enum qk_multi {}

extern "C" {
    pub fn qk_abi_version() -> i32;
    pub fn qk_build_id() -> *const c_char;
    pub fn qk_last_error() -> *const c_char;

    pub fn qk_model_create(out: *mut *mut qk_model) -> i32;
    pub fn qk_model_destroy(m: *mut qk_model);
    pub fn qk_model_add_component(m: *mut qk_model, d: *const qk_component_desc, out_id: *mut u32) -> i32;
    pub fn qk_model_set_dist(m: *mut qk_model, comp: u32, which: u32, d: *const qk_dist_desc, out_id: *mut u32) -> i32;
    pub fn qk_model_add_delay_mode(
        m: *mut qk_model,
        comp: u32,
        until_delay: *const qk_dist_desc,
        until_fix: *const qk_dist_desc,
        out_until_delay: *mut u32,
        out_until_fix: *mut u32,
    ) -> i32;
    pub fn qk_model_set_vector(m: *mut qk_model, comp: u32, dim: u32, low: f64, max: f64, init3: *const f64) -> i32;
    /// user-defined resource types (iron_ore.rs:21-96): `dim` <= QK_VEC_MAX_DIM f64 fields, `total()` sums the fields of
    /// `total_mask` in field order, `remove(q)` takes q / total() of every field; `init[dim]`
    pub fn qk_model_set_vector_n(m: *mut qk_model, comp: u32, dim: u32, total_mask: u32, low: f64, max: f64, init: *const f64) -> i32;
    pub fn qk_model_set_ports(m: *mut qk_model, comp: u32, n_ports: u32, ratios5: *const f64) -> i32;
    pub fn qk_model_connect(m: *mut qk_model, a: u32, b: u32, port_n: i32) -> i32;
    pub fn qk_model_set_logged(m: *mut qk_model, comp: u32, enabled: i32) -> i32;
    pub fn qk_model_schedule_env_state(m: *mut qk_model, t_ns: u64, env: u32, state: u32) -> i32;
    pub fn qk_model_schedule_low_capacity(m: *mut qk_model, t_ns: u64, stock: u32, value: f64) -> i32;
    /// create_scheduled_event!(SetProcessQuantity(cfg)) on an F64Process (core.rs:1387-1391)
    pub fn qk_model_schedule_process_quantity(m: *mut qk_model, t_ns: u64, process: u32, dist: *const qk_dist_desc, out_dist_id: *mut u32) -> i32;
    pub fn qk_model_n_components(m: *const qk_model, out: *mut u32) -> i32;
    pub fn qk_model_n_dists(m: *const qk_model, out: *mut u32) -> i32;
    pub fn qk_model_state_bytes(m: *const qk_model, with_log: i32, out_bytes: *mut u32) -> i32;
    /// Vector3ContainerFactory (vector_container.rs:126-156); `create_quantity` null = None
    pub fn qk_model_set_container_factory(
        m: *mut qk_model,
        source: u32,
        create_quantity: *const qk_dist_desc,
        split3: *const f64,
        capacity: f64,
        out_dist_id: *mut u32,
    ) -> i32;
    /// with_initial_resource(ItemDeque<Vector3Container>) of a container stock
    pub fn qk_model_set_initial_containers(
        m: *mut qk_model,
        stock: u32,
        n: u32,
        resources: *const f64,
        has_resource: *const u8,
        capacities: *const f64,
    ) -> i32;

    pub fn qk_batch_create(m: *const qk_model, cfg: *const qk_batch_config, out: *mut *mut qk_batch) -> i32;
    pub fn qk_batch_info_get(b: *mut qk_batch, out: *mut qk_batch_info) -> i32;
    pub fn qk_batch_destroy(b: *mut qk_batch);
    pub fn qk_batch_set_tape(b: *mut qk_batch, dist_id: u32, values: *const f64, per_rep_len: u64) -> i32;
    pub fn qk_batch_init(b: *mut qk_batch, t0_ns: u64) -> i32;
    pub fn qk_batch_step_until(b: *mut qk_batch, t_ns: u64) -> i32;
    pub fn qk_batch_step_events(b: *mut qk_batch, n_events: u64) -> i32;
    pub fn qk_batch_step_events_chunked(b: *mut qk_batch, n_events: u64, chunk: u64) -> i32;
    pub fn qk_batch_step_events_sliced(b: *mut qk_batch, n_events: u64, n_chunks: u32) -> i32;
    pub fn qk_batch_synchronize(b: *mut qk_batch) -> i32;
    pub fn qk_batch_summary_get(b: *mut qk_batch, out: *mut qk_batch_summary) -> i32;
    pub fn qk_batch_reduce_stats(b: *mut qk_batch, out: *mut qk_comp_summary, n_comp: u32) -> i32;
    pub fn qk_batch_reduce_partials(b: *mut qk_batch, partials: *mut u64, n_comp: u32) -> i32;
    pub fn qk_batch_reduce_partials_device(b: *mut qk_batch, d_partials: *mut u64, n_comp: u32) -> i32;
    pub fn qk_batch_combine_partials_device(b: *mut qk_batch, d_parts: *const u64, n_parts: u32, d_out: *mut u64, n_comp: u32) -> i32;
    pub fn qk_batch_histograms_device(b: *mut qk_batch, d_partials: *const u64, d_hist: *mut u64, n_comp: u32) -> i32;
    pub fn qk_batch_read_logs(b: *mut qk_batch, rep0: u64, n: u64, buf: *mut qk_log_record, kept: *mut u32, total: *mut u64) -> i32;
    pub fn qk_batch_read_vector_n(b: *mut qk_batch, rep: u64, comp: u32, out: *mut f64, cap: u32, dim_out: *mut u32) -> i32;
    pub fn qk_batch_last_step_ms(b: *mut qk_batch, ms: *mut f32, n_launches: *mut u32) -> i32;

    // one job on every GPU of the box: shards by contiguous replication range, NCCL only inside qk_multi_stats
    pub fn qk_multi_create(m: *const qk_model, cfg: *const qk_multi_config, out: *mut *mut qk_multi) -> i32;
    pub fn qk_multi_destroy(j: *mut qk_multi);
    pub fn qk_multi_n_shards(j: *mut qk_multi, n: *mut u32) -> i32;
    pub fn qk_multi_shard(j: *mut qk_multi, i: u32, batch: *mut *mut qk_batch, rep0: *mut u64, n_reps: *mut u64, device: *mut i32) -> i32;
    pub fn qk_multi_set_tape(j: *mut qk_multi, dist_id: u32, values: *const f64, per_rep_len: u64) -> i32;
    pub fn qk_multi_init(j: *mut qk_multi, t0_ns: u64) -> i32;
    pub fn qk_multi_step_until(j: *mut qk_multi, t_ns: u64) -> i32;
    pub fn qk_multi_step_events(j: *mut qk_multi, n_events: u64) -> i32;
    pub fn qk_multi_synchronize(j: *mut qk_multi) -> i32;
    pub fn qk_multi_last_step_ms(j: *mut qk_multi, ms_max: *mut f32) -> i32;
    pub fn qk_multi_stats(j: *mut qk_multi, out: *mut qk_comp_summary, n_comp: u32, total: *mut qk_batch_summary) -> i32;
    pub fn qk_batch_read_status(b: *mut qk_batch, rep0: u64, n: u64, status: *mut u32, now_ns: *mut u64, n_events: *mut u64) -> i32;
    pub fn qk_batch_comp_stats(b: *mut qk_batch, rep0: u64, n: u64, comp: u32, out: *mut qk_comp_stats) -> i32;
    pub fn qk_batch_read_log(b: *mut qk_batch, rep: u64, buf: *mut qk_log_record, cap: u64, n_out: *mut u64, n_total: *mut u64) -> i32;
    pub fn qk_batch_read_stock(b: *mut qk_batch, rep: u64, comp: u32, items: *mut u32, cap: u64, n_out: *mut u64) -> i32;
    pub fn qk_batch_read_vector(b: *mut qk_batch, rep: u64, comp: u32, out3: *mut f64) -> i32;
    pub fn qk_batch_read_containers(
        b: *mut qk_batch,
        rep: u64,
        comp: u32,
        handles: *mut u32,
        resources: *mut u64,
        cap: u64,
        n_out: *mut u64,
    ) -> i32;
}

/// `Err(String)` carrying the library's thread-local message, like the reference's `Box<dyn Error>` strings.
pub fn check(rc: i32) -> Result<(), Box<dyn std::error::Error>> {
    if rc == QK_OK {
        return Ok(());
    }
    let msg = unsafe {
        let p = qk_last_error();
        if p.is_null() {
            format!("libquokka_b200: error {}", rc)
        } else {
            std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    };
    Err(msg.into())
}
