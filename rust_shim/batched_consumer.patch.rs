// Batched shim for PairEndScanner (src/core/pescanner.rs:350-518).  SOURCE ONLY (no Rust toolchain in the build image).
//
// Upstream's consumers pop a pack of 1000 pairs (src/core/common.rs:23) and call scan_pair_end(pack), which calls
// FusionMapper::map_read 1..4 times per pair.  Mapping one pack per gf_map_pairs call costs ~70 us of launches and
// synchronisation per 1000 pairs (profiles/r02_small_batch.jsonl); the aggregator below hands the packs to a
// gf_stream (include/genefuse_gpu.h), which copies the strings into pinned arenas and maps >= 2^20 pairs per call.
//
// What changes in PairEndScanner:
//   * new fields      gpu_stream: PackStream,  held: Mutex<HashMap<u64, ReadPairPack<'s>>>   (packs whose pairs may still match)
//   * ReadPairPack    gets `first_pair: u64` = running pair count of the producer when the pack was cut
//                     (producer_task, pescanner.rs:190-249: `first_pair += pack.count` after each push)
//   * consume_pack    calls push_pack(pack) instead of scan_pair_end(pack)
//   * _scan           after the producer / consumers have joined (pescanner.rs:296-311): finish_stream()
// Everything downstream (push_match -> add_match buckets, filter / sort / cluster / report) is untouched.

fn push_pack(&self, pack: ReadPairPack<'s>) -> Result<(), Error> {
    let n = pack.count as usize;
    let (mut s1, mut q1, mut s2, mut q2) = (Vec::with_capacity(n), Vec::with_capacity(n), Vec::with_capacity(n), Vec::with_capacity(n));
    let (mut l1, mut l2) = (Vec::with_capacity(n), Vec::with_capacity(n));
    for pair in pack.data.iter().take(n) {
        s1.push(pair.m_left.m_seq.m_str.as_ptr());   q1.push(pair.m_left.m_quality.as_ptr());   l1.push(pair.m_left.len() as u32);
        s2.push(pair.m_right.m_seq.m_str.as_ptr());  q2.push(pair.m_right.m_quality.as_ptr());  l2.push(pair.m_right.len() as u32);
    }
    // the library copies the bytes before it returns: the strings are only borrowed for the call
    self.gpu_stream.push(pack.first_pair, &s1, &q1, &l1, &s2, &q2, &l2).map_err(|e| -> Error { e.into() })?;
    self.held.lock().unwrap().insert(pack.first_pair / PACK_SIZE as u64, pack);
    self.drain_records()
}

/// records of the batches the stream has mapped so far -> ReadMatch -> push_match (same rebuild as scan_pair_end.patch.rs)
fn drain_records(&self) -> Result<(), Error> {
    for m in self.gpu_stream.take().map_err(|e| -> Error { e.into() })? {
        let held = self.held.lock().unwrap();
        let pack = held.get(&(m.pair_idx / PACK_SIZE as u64)).unwrap();
        let pair = &pack.data[(m.pair_idx - pack.first_pair) as usize];
        let mut read = match m.source { 0 => pair.fast_merge().unwrap(), 1 => pair.m_left.clone(), _ => pair.m_right.clone() };
        if m.used_rc != 0 { read = read.reverse_complement(); }
        let mut rm = ReadMatch::new(read, m.read_break, GenePos { contig: m.l_contig, position: m.l_pos },
                                    GenePos { contig: m.r_contig, position: m.r_pos }, m.gap, false);
        rm.m_left_distance = m.l_dist;
        rm.m_right_distance = m.r_dist;
        rm.add_original_pair(pair.clone());
        if m.reversed != 0 { rm.set_reversed(true); }
        drop(held);
        self.push_match(rm);
    }
    // packs older than the stream's current batch can no longer produce records: release them
    let done_before = self.gpu_stream.pairs_mapped() / PACK_SIZE as u64;
    self.held.lock().unwrap().retain(|k, _| *k >= done_before);
    Ok(())
}

fn finish_stream(&self) -> Result<(), Error> {
    self.gpu_stream.flush().map_err(|e| -> Error { e.into() })?;
    self.drain_records()
}

// ---- FusionMapper::filter_matches (src/core/fusion_mapper.rs:276-296) with the device-side stages ------------------
// remove_by_complexity / remove_by_distance / remove_indels: `gf_index_set_output_mode(h, GF_OUT_DROP_FILTERED)` before the scan
// makes the library drop those records before they leave the device (their predicates are evaluated in k_verify); the three
// retain() passes then find nothing to remove and can stay as they are.
//
// remove_alignables (src/core/fusion_mapper.rs:488-542): replace Matcher::from_ref_and_seqs + the do_match loop by
//     let reference = GpuReference::build(&contigs_in_name_order, device)?;      // once per FASTA; list mode: once for all CSVs
//     let res = reference.alignable_filter(&seqs)?;                              // seqs gathered exactly as at :496-500
//     if res.panic_stage != 0 { panic!("Matcher would panic here (stage {}, sequence {})", res.panic_stage, res.panic_seq); }
//     // res.n_removed == 0 always: nothing is retained away (include/genefuse_gpu.h explains why)
// This removes the 13-18 s whole-genome pass of the reference (benchmark_res/bench_res.md:8-9) from every run.
//
// cluster_matches (src/core/fusion_mapper.rs:394-486): see GpuIndex::adjust_fusion_break in genefuse_gpu.rs.
