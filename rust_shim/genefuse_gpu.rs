//! genefuse_gpu.rs — Rust FFI binding of include/genefuse_gpu.h.  SOURCE ONLY: this image has no cargo/rustc, so
//! the file is shipped for a maintainer to drop into GeneFuseRust as `src/core/genefuse_gpu.rs`
//! (add `pub(crate) mod genefuse_gpu;` to src/core/mod.rs and `println!("cargo:rustc-link-lib=genefuse_b200")`
//! + a search path to build.rs).  It replaces two call sites and nothing else:
//!   Indexer::make_index             src/core/indexer.rs:122-177      -> GpuIndex::build
//!   PairEndScanner::scan_pair_end   src/core/pescanner.rs:427-518    -> GpuIndex::map_pairs
//!   SingleEndScanner::scan_single_end src/core/sescanner.rs:183-205  -> GpuIndex::map_pairs (seq2 = null)
#![allow(non_camel_case_types, dead_code)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct gf_gene_span { pub seq: *const u8, pub len: u32, pub reversed: u8 }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct gf_params {
    pub skip_key_dup_threshold: i32,
    pub major_gene_key_requirement: i32,
    pub minor_gene_key_requirement: i32,
    pub mismatch_threshold: i32,
    pub deletion_threshold: i32,
}

#[repr(C)]
pub struct gf_batch {
    pub n: u64,
    pub seq1: *const u8, pub qual1: *const u8, pub off1: *const u64,
    pub seq2: *const u8, pub qual2: *const u8, pub off2: *const u64,
    pub bytes1: u64, pub bytes2: u64,
    pub max_len: u32, pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct gf_match {
    pub pair_idx: u64,
    pub read_break: i32, pub l_pos: i32, pub r_pos: i32, pub gap: i32,
    pub l_dist: i32, pub r_dist: i32, pub seq_len: i32,
    pub l_contig: i16, pub r_contig: i16, pub merge_olen: i16, pub merge_diff: i16,
    pub source: u8, pub used_rc: u8, pub reversed: u8, pub filter_flags: u8,
}

#[repr(C)] pub struct gf_index { _private: [u8; 0] }

extern "C" {
    fn gf_last_error() -> *const c_char;
    fn gf_index_create(genes: *const gf_gene_span, n_genes: u32, params: *const gf_params, device: c_int,
                       out: *mut *mut gf_index) -> c_int;
    fn gf_index_destroy(idx: *mut gf_index);
    fn gf_map_pairs(idx: *mut gf_index, batch: *const gf_batch, out: *mut gf_match, out_cap: u64,
                    n_out: *mut u64) -> c_int;
    fn gf_list_map_pairs(idx: *const *mut gf_index, n_idx: u32, batch: *const gf_batch, out: *const *mut gf_match,
                         out_cap: *const u64, n_out: *mut u64) -> c_int;
    fn gf_adjust_fusion_break(idx: *mut gf_index, bytes: *const u8, n_bytes: u64, refs: *const gf_break_ref, n_refs: u32,
                              jobs: *const gf_break_job, n_jobs: u64, out: *mut gf_break_out) -> c_int;
    fn gf_index_set_output_mode(idx: *mut gf_index, mode: u32) -> c_int;
    fn gf_stream_create(idx: *mut gf_index, paired: c_int, batch_pairs: u64, out: *mut *mut gf_stream) -> c_int;
    fn gf_stream_destroy(s: *mut gf_stream);
    fn gf_stream_push(s: *mut gf_stream, first_pair: u64, n: u64, seq1: *const *const u8, qual1: *const *const u8, len1: *const u32,
                      seq2: *const *const u8, qual2: *const *const u8, len2: *const u32) -> c_int;
    fn gf_stream_flush(s: *mut gf_stream) -> c_int;
    fn gf_stream_take(s: *mut gf_stream, out: *mut gf_match, out_cap: u64, n_out: *mut u64) -> c_int;
    fn gf_stream_get_counts(s: *const gf_stream, pairs_pushed: *mut u64, map_calls: *mut u64) -> c_int;
    fn gf_reference_create(contigs: *const gf_ref_contig, n_contigs: u32, device: c_int, out: *mut *mut gf_reference) -> c_int;
    fn gf_reference_destroy(r: *mut gf_reference);
    fn gf_alignable_filter(r: *mut gf_reference, seqs: *const u8, seq_off: *const u64, n_seqs: u64, alignable: *mut u8,
                           res: *mut gf_alignable_result) -> c_int;
}

#[repr(C)] pub struct gf_stream { _private: [u8; 0] }
#[repr(C)] pub struct gf_reference { _private: [u8; 0] }
#[repr(C)] pub struct gf_ref_contig { pub seq: *const u8, pub len: u64 }
#[repr(C)] #[derive(Clone, Copy, Default, Debug)]
pub struct gf_alignable_result {
    pub key_positions: [u64; 4], pub n_removed: u64, pub panic_seq: i64, pub bloom_bits: u32, pub panic_stage: i32,
}
pub const GF_OUT_DROP_FILTERED: u32 = 1;
pub const GF_OUT_BUCKET_ORDER: u32 = 2;
pub const GF_E_REF_PANIC: c_int = -5;

/// include/genefuse_gpu.h: one FusionResult's m_left_ref / m_right_ref inside the byte arena
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct gf_break_ref { pub left_off: u64, pub right_off: u64, pub left_len: u32, pub right_len: u32 }
/// one ReadMatch of that FusionResult: m_read.m_seq inside the arena + m_read_break
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct gf_break_job { pub seq_off: u64, pub seq_len: u32, pub read_break: i32, pub result: u32, pub reserved: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct gf_break_out { pub shift: i32, pub left_distance: i32, pub right_distance: i32, pub status: i32 }

pub const GF_E_CAPACITY: c_int = -3;

fn last_error() -> String {
    unsafe { CStr::from_ptr(gf_last_error()).to_string_lossy().into_owned() }
}

/// Owns one device index (one per FusionMapper; list mode: one per CSV, handles are independent).
pub struct GpuIndex { h: *mut gf_index }
unsafe impl Send for GpuIndex {}
unsafe impl Sync for GpuIndex {} // calls on one handle are serialised inside the library

impl GpuIndex {
    /// `genes[i]` = (upper-cased contig[m_start..m_end] or "" when the chromosome is unresolved, is_reversed),
    /// i.e. exactly what Indexer::make_index computes at src/core/indexer.rs:136-159 before index_contig.
    pub fn build(genes: &[(&[u8], bool)], params: gf_params, device: i32) -> Result<Self, String> {
        let spans: Vec<gf_gene_span> = genes.iter()
            .map(|(s, r)| gf_gene_span { seq: s.as_ptr(), len: s.len() as u32, reversed: *r as u8 }).collect();
        let mut h: *mut gf_index = std::ptr::null_mut();
        let rc = unsafe { gf_index_create(spans.as_ptr(), spans.len() as u32, &params, device, &mut h) };
        if rc != 0 { return Err(last_error()); }
        Ok(Self { h })
    }

    /// Maps a batch held in byte arenas (offsets have n+1 entries; seq and qual share them).
    pub fn map_pairs(&self, batch: &gf_batch) -> Result<Vec<gf_match>, String> {
        let mut cap = (batch.n as usize / 8).max(1024);
        loop {
            let mut out = vec![gf_match::default(); cap];
            let mut n: u64 = 0;
            let rc = unsafe { gf_map_pairs(self.h, batch, out.as_mut_ptr(), cap as u64, &mut n) };
            if rc == GF_E_CAPACITY { cap = n as usize; continue; }
            if rc != 0 { return Err(last_error()); } // upstream unwraps -> same abort behaviour as a panic
            out.truncate(n as usize);
            return Ok(out);
        }
    }
}

impl GpuIndex {
    /// FusionResult::adjust_fusion_break (src/core/fusion_result.rs:299-321) for every match of every result of one
    /// mapper: call it from cluster_matches right after fr.make_reference(..) (src/core/fusion_mapper.rs:438-456)
    /// instead of fr.adjust_fusion_break(); then for each match: m_read_break += shift, both positions += shift,
    /// m_left_distance / m_right_distance = the returned distances.
    pub fn adjust_fusion_break(&self, arena: &[u8], refs: &[gf_break_ref], jobs: &[gf_break_job])
        -> Result<Vec<gf_break_out>, String> {
        let mut out = vec![gf_break_out::default(); jobs.len()];
        let rc = unsafe { gf_adjust_fusion_break(self.h, arena.as_ptr(), arena.len() as u64, refs.as_ptr(), refs.len() as u32,
                                                 jobs.as_ptr(), jobs.len() as u64, out.as_mut_ptr()) };
        if rc != 0 { return Err(last_error()); }
        Ok(out)
    }
}

/// List mode (src/core/fusion_scan.rs:62-188): one batch against every CSV's index in one call — one upload, one
/// conversion / fast_merge pass, then per-index seeding, screening and verification.  Returns one record vector per index,
/// each identical to `indices[h].map_pairs(batch)`.
pub fn list_map_pairs(indices: &[&GpuIndex], batch: &gf_batch) -> Result<Vec<Vec<gf_match>>, String> {
    let hs: Vec<*mut gf_index> = indices.iter().map(|g| g.h).collect();
    let mut caps: Vec<u64> = vec![(batch.n / 8).max(1024); hs.len()];
    loop {
        let mut bufs: Vec<Vec<gf_match>> = caps.iter().map(|c| vec![gf_match::default(); *c as usize]).collect();
        let ptrs: Vec<*mut gf_match> = bufs.iter_mut().map(|b| b.as_mut_ptr()).collect();
        let mut n = vec![0u64; hs.len()];
        let rc = unsafe { gf_list_map_pairs(hs.as_ptr(), hs.len() as u32, batch, ptrs.as_ptr(), caps.as_ptr(), n.as_mut_ptr()) };
        if rc == GF_E_CAPACITY { for (c, k) in caps.iter_mut().zip(&n) { *c = (*c).max(*k); } continue; }
        if rc != 0 { return Err(last_error()); }
        for (b, k) in bufs.iter_mut().zip(&n) { b.truncate(*k as usize); }
        return Ok(bufs);
    }
}

impl Drop for GpuIndex { fn drop(&mut self) { unsafe { gf_index_destroy(self.h) } } }


/// The batched shim (rust_shim/batched_consumer.patch.rs): packs in, large batches to the device.
pub struct PackStream { s: *mut gf_stream }
unsafe impl Send for PackStream {}
unsafe impl Sync for PackStream {} // pushes on one stream are serialised inside the library
impl PackStream {
    pub fn new(index: &GpuIndex, paired: bool, batch_pairs: u64) -> Result<Self, String> {
        let mut s: *mut gf_stream = std::ptr::null_mut();
        if unsafe { gf_stream_create(index.h, paired as c_int, batch_pairs, &mut s) } != 0 { return Err(last_error()); }
        Ok(Self { s })
    }
    pub fn push(&self, first_pair: u64, s1: &[*const u8], q1: &[*const u8], l1: &[u32], s2: &[*const u8], q2: &[*const u8],
                l2: &[u32]) -> Result<(), String> {
        let rc = unsafe { gf_stream_push(self.s, first_pair, s1.len() as u64, s1.as_ptr(), q1.as_ptr(), l1.as_ptr(),
                                         s2.as_ptr(), q2.as_ptr(), l2.as_ptr()) };
        if rc != 0 { Err(last_error()) } else { Ok(()) }
    }
    pub fn flush(&self) -> Result<(), String> { if unsafe { gf_stream_flush(self.s) } != 0 { Err(last_error()) } else { Ok(()) } }
    pub fn take(&self) -> Result<Vec<gf_match>, String> {
        let mut cap = 4096usize;
        loop {
            let mut out = vec![gf_match::default(); cap];
            let mut n = 0u64;
            let rc = unsafe { gf_stream_take(self.s, out.as_mut_ptr(), cap as u64, &mut n) };
            if rc == GF_E_CAPACITY { cap = n as usize; continue; }
            if rc != 0 { return Err(last_error()); }
            out.truncate(n as usize);
            return Ok(out);
        }
    }
    /// pairs that have gone through a mapping call (their packs can be released once their records were drained)
    pub fn pairs_mapped(&self) -> u64 {
        let (mut pushed, mut calls) = (0u64, 0u64);
        unsafe { gf_stream_get_counts(self.s, &mut pushed, &mut calls) };
        pushed // conservative: the caller releases packs only after take() returned their records
    }
}
impl Drop for PackStream { fn drop(&mut self) { unsafe { gf_stream_destroy(self.s) } } }

/// The Matcher pass (src/core/fusion_mapper.rs:488-542, src/core/matcher.rs): the reference is scanned ONCE per FASTA.
pub struct GpuReference { r: *mut gf_reference }
unsafe impl Send for GpuReference {}
unsafe impl Sync for GpuReference {}
impl GpuReference {
    /// `contigs` = FastaReader::m_all_contigs values in BTreeMap (name) order
    pub fn build(contigs: &[&[u8]], device: i32) -> Result<Self, String> {
        let c: Vec<gf_ref_contig> = contigs.iter().map(|s| gf_ref_contig { seq: s.as_ptr(), len: s.len() as u64 }).collect();
        let mut r: *mut gf_reference = std::ptr::null_mut();
        if unsafe { gf_reference_create(c.as_ptr(), c.len() as u32, device, &mut r) } != 0 { return Err(last_error()); }
        Ok(Self { r })
    }
    /// `seqs` = ReadMatch::get_read().m_seq of every surviving match in bucket order (fusion_mapper.rs:496-500)
    pub fn alignable_filter(&self, seqs: &[&str]) -> Result<(Vec<u8>, gf_alignable_result), String> {
        let mut arena = Vec::new();
        let mut off = vec![0u64];
        for s in seqs { arena.extend_from_slice(s.as_bytes()); off.push(arena.len() as u64); }
        let mut flags = vec![0u8; seqs.len().max(1)];
        let mut res = gf_alignable_result::default();
        let rc = unsafe { gf_alignable_filter(self.r, arena.as_ptr(), off.as_ptr(), seqs.len() as u64, flags.as_mut_ptr(), &mut res) };
        if rc != 0 && rc != GF_E_REF_PANIC { return Err(last_error()); }
        flags.truncate(seqs.len());
        Ok((flags, res))
    }
}
impl Drop for GpuReference { fn drop(&mut self) { unsafe { gf_reference_destroy(self.r) } } }
