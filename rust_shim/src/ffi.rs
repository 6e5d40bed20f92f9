//! The `extern "C"` block of include/pcp_b200.h, entry point by entry point (NOT COMPILED HERE:
//! no Rust toolchain in this image).  Field order and widths follow the header; tests/test_abi.py
//! pins the struct sizes the C side has (Operand 8, Config 16, Stats 24, SearchConfig 40,
//! SearchResult 80 bytes).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct PcpOperand {
    pub var: i32,
    pub off: i32,
}

#[repr(C)]
pub struct PcpConfig {
    pub device: i32,
    pub flags: u32,
    pub max_labels: u32,
    pub tail_limit: u32,
}

#[repr(C)]
#[derive(Default)]
pub struct PcpStats {
    pub propagations: u64,
    pub iterations: u32,
    pub active_props: u32,
    pub kernel_ms: f32,
    pub launches: u32,
}

#[repr(C)]
#[derive(Default)]
pub struct PcpSearchConfig {
    pub node_limit: u64,
    pub all_solutions: i32,
    pub var_sel: i32,
    pub val_sel: i32,
    pub distributor: i32,
    pub bb_mode: i32,
    pub bb_var: i32,
    pub trace_domains: i32,
    pub warmup_nodes: i32,
}

#[repr(C)]
#[derive(Default)]
pub struct PcpSearchResult {
    pub status: i32,
    pub has_bb_value: i32,
    pub bb_value: i32,
    pub reserved: i32,
    pub num_nodes: u64,
    pub num_solution: u64,
    pub num_failed_node: u64,
    pub num_prune: u64,
    pub propagations: u64,
    pub iterations: u64,
    pub seconds: f64,
    pub kernel_seconds: f64,
}

pub enum PcpEngine {}
pub enum PcpSearch {}

pub const PCP_OK: c_int = 0;
pub const PCP_ERR_INVALID: c_int = -1;
pub const PCP_ERR_CUDA: c_int = -2;
pub const PCP_ERR_NOMEM: c_int = -3;
pub const PCP_ERR_UNSUPPORTED: c_int = -4;

pub const PCP_FALSE: i32 = -1;
pub const PCP_UNKNOWN: i32 = 0;
pub const PCP_TRUE: i32 = 1;

pub const PCP_VAR_CONSTANT: i32 = -1;
pub fn var_sum(sum_id: i32) -> i32 {
    -2 - sum_id
}

pub const PCP_X_LESS_Y: i32 = 0;
pub const PCP_X_NEQ_Y: i32 = 1;
pub const PCP_X_EQ_Y: i32 = 2;
pub const PCP_X_GREATER_Y_PLUS_Z: i32 = 3;
pub const PCP_X_LESS_Y_PLUS_Z: i32 = 4;
pub const PCP_X_EQ_Y_PLUS_Z: i32 = 5;
pub const PCP_DISTINCT: i32 = 6;
pub const PCP_DISJ2_X_EQ_Y_PLUS_Z: i32 = 7;
pub const PCP_X_EQ_Y_MUL_Z: i32 = 8;
pub const PCP_ALL_EQUAL: i32 = 9;

pub const PCP_F_CONJUNCTION: i32 = 1;
pub const PCP_F_DISJUNCTION: i32 = 2;
pub const PCP_F_BOOLEAN: i32 = 3;
pub const PCP_F_BOOLEAN_NEG: i32 = 4;
pub const PCP_F_NOT: i32 = 5;
pub const PCP_F_LEAF: i32 = 16;

pub const PCP_FLAG_INCREMENTAL: u32 = 1;
pub const PCP_FLAG_HOST_SEARCH: u32 = 2;
pub const PCP_FLAG_INTERVAL_SET: u32 = 4;

#[link(name = "pcp_b200")]
extern "C" {
    pub fn pcp_engine_create(cfg: *const PcpConfig, out: *mut *mut PcpEngine) -> c_int;
    pub fn pcp_engine_fork(parent: *mut PcpEngine, out: *mut *mut PcpEngine) -> c_int;
    pub fn pcp_engine_destroy(e: *mut PcpEngine);
    pub fn pcp_last_error(e: *const PcpEngine) -> *const c_char;
    pub fn pcp_set_timing(e: *mut PcpEngine, enabled: i32) -> c_int;
    pub fn pcp_set_grid_limit(e: *mut PcpEngine, max_ctas: i32) -> c_int;
    pub fn pcp_stream(e: *mut PcpEngine, stream: *mut *mut c_void) -> c_int;
    pub fn pcp_vars_alloc(e: *mut PcpEngine, lo: *const i32, hi: *const i32, n: i32, first_idx: *mut i32) -> c_int;
    pub fn pcp_sum_alloc(e: *mut PcpEngine, terms: *const PcpOperand, n: i32, sum_id: *mut i32) -> c_int;
    pub fn pcp_prop_alloc(e: *mut PcpEngine, kind: i32, ops: *const PcpOperand, n_ops: i32, idx: *mut i32) -> c_int;
    pub fn pcp_props_alloc(e: *mut PcpEngine, kind: i32, ops: *const PcpOperand, n_ops: i32, n_props: i64, first_idx: *mut i32) -> c_int;
    pub fn pcp_formula_alloc(e: *mut PcpEngine, words: *const i32, n_words: i32, idx: *mut i32) -> c_int;
    pub fn pcp_consistency(e: *mut PcpEngine, status: *mut i32, stats: *mut PcpStats) -> c_int;
    pub fn pcp_consistency_batch(engines: *const *mut PcpEngine, n: i32, status: *mut i32, stats: *mut PcpStats) -> c_int;
    pub fn pcp_domains_read(e: *mut PcpEngine, first: i32, n: i32, lo: *mut i32, hi: *mut i32) -> c_int;
    pub fn pcp_domains_size_read(e: *mut PcpEngine, first: i32, n: i32, size: *mut u32) -> c_int;
    pub fn pcp_domains_read_bits(e: *mut PcpEngine, first: i32, n: i32, base: i32, words: i32, out: *mut u32) -> c_int;
    pub fn pcp_var_update(e: *mut PcpEngine, idx: i32, lo: i32, hi: i32, ok: *mut i32) -> c_int;
    pub fn pcp_active_read(e: *mut PcpEngine, first: i32, n: i32, out: *mut u8) -> c_int;
    pub fn pcp_label(e: *mut PcpEngine, label: *mut u64) -> c_int;
    pub fn pcp_restore(e: *mut PcpEngine, label: u64) -> c_int;
    pub fn pcp_num_vars(e: *const PcpEngine, n: *mut i32) -> c_int;
    pub fn pcp_num_props(e: *const PcpEngine, n: *mut i32) -> c_int;
    pub fn pcp_search_run(e: *mut PcpEngine, cfg: *const PcpSearchConfig, res: *mut PcpSearchResult, trace_status: *mut i32,
                          trace_hash: *mut u64, trace_lo: *mut i32, trace_hi: *mut i32, trace_capacity: u64) -> c_int;
    pub fn pcp_search_open(e: *mut PcpEngine, cfg: *const PcpSearchConfig, trace_status: *mut i32, trace_hash: *mut u64,
                           trace_lo: *mut i32, trace_hi: *mut i32, trace_capacity: u64, out: *mut *mut PcpSearch) -> c_int;
    pub fn pcp_search_step(s: *mut PcpSearch, max_nodes: u64, res: *mut PcpSearchResult) -> c_int;
    pub fn pcp_search_step_many(searches: *const *mut PcpSearch, n: i32, max_nodes: u64, res: *mut PcpSearchResult) -> c_int;
    pub fn pcp_search_set_incumbent(s: *mut PcpSearch, value: i32) -> c_int;
    pub fn pcp_search_close(s: *mut PcpSearch);
}
