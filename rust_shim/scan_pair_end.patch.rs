// Replacement body for PairEndScanner::scan_pair_end (src/core/pescanner.rs:427-518).  SOURCE ONLY (no Rust
// toolchain in the build image).  The per-pair decision tree (merge -> map -> rc retry -> push) runs on the GPU;
// the host only rebuilds the strings of the few matched pairs and pushes them, so filter/sort/cluster/report
// code downstream is untouched.
fn scan_pair_end(&self, pack: ReadPairPack<'s>) -> Result<bool, Error> {
    let mapper = self.m_fusion_mapper_o.as_ref().unwrap();
    // 1. flatten the pack into arenas (the shim may also aggregate many packs into one >= 2^20-pair batch)
    let (mut s1, mut q1, mut s2, mut q2) = (Vec::new(), Vec::new(), Vec::new(), Vec::new());
    let (mut o1, mut o2) = (vec![0u64], vec![0u64]);
    let mut max_len = 0usize;
    for pair in pack.data.iter().take(pack.count as usize) {
        s1.extend_from_slice(pair.m_left.m_seq.m_str.as_bytes());
        q1.extend_from_slice(pair.m_left.m_quality.as_bytes());
        s2.extend_from_slice(pair.m_right.m_seq.m_str.as_bytes());
        q2.extend_from_slice(pair.m_right.m_quality.as_bytes());
        o1.push(s1.len() as u64);
        o2.push(s2.len() as u64);
        max_len = max_len.max(pair.m_left.len()).max(pair.m_right.len());
    }
    let batch = gf_batch { n: pack.count as u64, seq1: s1.as_ptr(), qual1: q1.as_ptr(), off1: o1.as_ptr(),
                           seq2: s2.as_ptr(), qual2: q2.as_ptr(), off2: o2.as_ptr(),
                           bytes1: s1.len() as u64, bytes2: s2.len() as u64, max_len: max_len as u32, reserved: 0 };
    // 2. one call replaces 1..4 FusionMapper::map_read calls per pair
    let records = mapper.m_indexer.gpu.map_pairs(&batch).map_err(|e| -> Error { e.into() })?;
    // 3. rebuild ReadMatch for the matched pairs only (records are sorted by (pair_idx, source))
    for m in records {
        let pair = &pack.data[m.pair_idx as usize];
        let mut read = match m.source {
            0 => pair.fast_merge().unwrap(),          // same merge the device found (olen = m.merge_olen)
            1 => pair.m_left.clone(),
            _ => pair.m_right.clone(),
        };
        if m.used_rc != 0 { read = read.reverse_complement(); }
        let mut rm = ReadMatch::new(read, m.read_break,
            GenePos { contig: m.l_contig, position: m.l_pos }, GenePos { contig: m.r_contig, position: m.r_pos },
            m.gap, false);
        rm.m_left_distance = m.l_dist;
        rm.m_right_distance = m.r_dist;
        rm.add_original_pair(pair.clone());
        if m.reversed != 0 { rm.set_reversed(true); }   // R1/R2 rc matches only (pescanner.rs:483,506)
        self.push_match(rm);
    }
    Ok(true)
}

// Indexer::make_index (src/core/indexer.rs:122-177): keep :136-159 (name resolution, slice, to_uppercase, push
// into m_fusion_seq), drop the two index_contig calls and fill_bloom_filter, and after the loop:
//     let genes: Vec<(&[u8], bool)> = self.m_fusion_seq.iter().zip(self.m_fusions.iter())
//         .map(|(s, f)| (s.as_bytes(), f.is_reversed())).collect();
//     self.gpu = GpuIndex::build(&genes, gf_params { skip_key_dup_threshold: gs.skip_key_dup_threshold as i32,
//         major_gene_key_requirement: gs.major_gene_key_requirement, minor_gene_key_requirement:
//         gs.minor_gene_key_requirement, mismatch_threshold: gs.mismatch_threshold,
//         deletion_threshold: gs.deletion_threshold as i32 }, device).unwrap();
// m_kmer_pos / m_dupe_list / m_bloom_filter (indexer.rs:74-76) and the 512 MiB allocation (:94,108) disappear.
