#!/usr/bin/env python3
"""Secondary configs of BASELINE.json (not the bench.py contract): the read-length sweep (config 5) and the
multi-CSV list mode (config 4), device-timed on one B200, each checked against the CPU oracle on a sample.

  python tools/bench_configs.py --sweep 75:50000000 250:50000000     # read_len:pairs
  python tools/bench_configs.py --list 16 --pairs 10000000
  python tools/bench_configs.py --fastq 4000000                      # raw FASTQ text, record splitting on the device
Prints one JSON line per config.

Developer tool (test / measurement infrastructure, not product code): the CPU oracle is loaded here only as the checker of
the CUDA path's records and as the reported CPU rate; the package under genefuserust_b200/ never touches it."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def device_batch(torch, synth, panel, P, L, seed, threads):
    pinned = [torch.empty(P * L, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
    batch = synth.generate_pairs(panel, P, read_len=L, seed=seed, threads=threads, out=tuple(t.numpy() for t in pinned))
    d = [t.cuda(non_blocking=True) for t in pinned]
    off = torch.arange(P + 1, dtype=torch.int64) * L
    d_off = off.cuda()
    torch.cuda.synchronize()
    return batch, d, d_off


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", nargs="*", default=[])
    ap.add_argument("--list", type=int, default=0)
    ap.add_argument("--fastq", type=int, default=0, help="pairs of raw FASTQ text through gf_map_fastq (device-side ingest)")
    ap.add_argument("--pairs", type=int, default=10_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--oracle-sample", type=int, default=300_000)
    a = ap.parse_args()
    import torch
    import __graft_entry__ as ge
    ge.build()
    import _oracle
    from genefuserust_b200 import synth, ReadBatch
    from genefuserust_b200._abi import gf_batch, gf_map_stats, gf_match
    from genefuserust_b200.host import FusionMapper
    threads = os.cpu_count() or 8
    panel = synth.make_panel(scale=1.0)
    genes = panel.genes()
    FusionMapper.from_gene_spans(genes[:4]).close()   # CUDA init / module load
    seeds = {75: 11, 150: 12, 250: 13}

    def run(mappers, d, d_off, P, L, steps):
        out_cap = max(1 << 16, P // 4)
        d_out = torch.empty(out_cap * C.sizeof(gf_match), dtype=torch.uint8, device="cuda")
        d_n = torch.zeros(1, dtype=torch.int64, device="cuda")
        db = gf_batch()
        db.n = P
        db.seq1, db.qual1, db.seq2, db.qual2 = (t.data_ptr() for t in d)
        db.off1 = db.off2 = d_off.data_ptr()
        db.bytes1 = db.bytes2 = P * L
        db.max_len = L
        st = torch.cuda.current_stream()
        lib = mappers[0].lib

        def one_pass():
            n = 0
            for m in mappers:
                rc = lib.gf_map_pairs_device(m.m_indexer.h, C.byref(db), d_out.data_ptr(), out_cap, d_n.data_ptr(),
                                             C.c_void_p(st.cuda_stream))
                assert rc == 0, lib.gf_last_error()
            return n
        for _ in range(3):
            one_pass()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(steps):
            one_pass()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        s = gf_map_stats()
        lib.gf_get_map_stats(mappers[-1].m_indexer.h, C.byref(s))
        return ms, s, int(d_n.item())

    for item in a.sweep:
        L, P = (int(x) for x in item.split(":"))
        batch, d, d_off = device_batch(torch, synth, panel, P, L, seeds.get(L, 12), threads)
        m = FusionMapper.from_gene_spans(genes)
        ms, s, n_matches = run([m], d, d_off, P, L, a.steps)
        # parity + CPU rate on a sample
        sample = batch.slice(0, min(a.oracle_sample, P))
        o = _oracle.OracleIndex(genes)
        t0 = time.perf_counter()
        want = o.scan(sample, threads=threads)
        cpu_dt = time.perf_counter() - t0
        got = [r.astuple() for r in m.scan_pair_end(sample)]
        alg = s.seq_bytes + 32 * s.n_probes_pass1
        print(json.dumps({"config": f"read-length sweep 2x{L}bp, {P} pairs, cancer-shaped panel, 1 B200", "pairs_per_s": P / (ms / 1e3),
                          "ms_per_step": ms, "screen_ms": s.ms_screen, "exact_verify_ms": s.ms_exact, "matches": n_matches,
                          "survivors": int(s.n_survivors), "roofline_frac_of_6538.3": alg / (s.ms_screen / 1e3) / 1e9 / 6538.3,
                          "parity_sample_pairs": sample.n, "parity_ok": got == want, "parity_matches": len(want),
                          "cpu_oracle_pairs_per_s": sample.n / cpu_dt, "cpu_cores": threads}), flush=True)
        assert got == want
        m.close()
        o.close()
        del d, d_off, batch
        torch.cuda.empty_cache()

    if a.fastq:
        # raw FASTQ text (pinned host memory) -> gf_map_fastq: H2D of the text, newline scan and record
        # tables on the device, mapping straight from the text.  Fixed-width records built with numpy.
        import numpy as np
        P, L = a.fastq, 150
        batch = synth.generate_pairs(panel, P, read_len=L, seed=12, threads=threads)

        def fastq_text(seq, qual, mate):
            name = np.frombuffer(b"@SYN:12:000000000 %d:N:0:ACGT\n" % mate, dtype=np.uint8)
            rec = len(name) + L + 1 + 2 + L + 1
            t = torch.empty(P * rec, dtype=torch.uint8, pin_memory=True)
            out = t.numpy().reshape(P, rec)
            out[:, :len(name)] = name
            idx = np.arange(P, dtype=np.int64)
            for k in range(9):                                   # zero-padded decimal pair index
                out[:, 8 + 8 - k] = 48 + (idx // 10 ** k) % 10
            o = len(name)
            out[:, o:o + L] = seq.reshape(P, L)
            out[:, o + L] = 10
            out[:, o + L + 1] = 43
            out[:, o + L + 2] = 10
            out[:, o + L + 3:o + 2 * L + 3] = qual.reshape(P, L)
            out[:, o + 2 * L + 3] = 10
            return t
        t1 = fastq_text(batch.seq1, batch.qual1, 1)
        t2 = fastq_text(batch.seq2, batch.qual2, 2)
        m = FusionMapper.from_gene_spans(genes)
        lib = m.lib
        cap = max(1 << 16, P // 4)
        out = (gf_match * cap)()
        n, nrec = C.c_uint64(0), C.c_uint64(0)

        def call():
            rc = lib.gf_map_fastq(m.m_indexer.h, C.cast(t1.data_ptr(), C.c_char_p), t1.numel(), C.cast(t2.data_ptr(), C.c_char_p),
                                  t2.numel(), out, cap, C.byref(n), C.byref(nrec))
            assert rc == 0, lib.gf_last_error()
        call()
        walls = []
        s = gf_map_stats()
        for _ in range(a.steps):
            t0 = time.perf_counter()
            call()
            walls.append(time.perf_counter() - t0)
            lib.gf_get_map_stats(m.m_indexer.h, C.byref(s))
            print(f"  call {len(walls)}: wall {walls[-1] * 1e3:.1f} ms, device total {s.ms_total:.1f} ms, ingest {s.ms_ingest:.1f} ms",
                  file=sys.stderr)
        dt = min(walls)
        got = [out[i].astuple() for i in range(n.value)]
        sample = batch.slice(0, min(a.oracle_sample, P))
        o = _oracle.OracleIndex(genes)
        want = o.scan(sample, threads=threads)
        ok = [g for g in got if g[0] < sample.n] == want
        text_bytes = t1.numel() + t2.numel()
        print(json.dumps({"config": f"FASTQ ingest on the device: {P} pairs 2x150 as raw text ({text_bytes} bytes, pinned host), 1 B200",
                          "records": int(nrec.value), "pairs_per_s_e2e": P / dt, "ms_per_call": dt * 1e3, "device_ms_total": s.ms_total,
                          "ingest_ms (H2D + newline scan + record tables)": s.ms_ingest,
                          "h2d_GBps_if_ingest_were_copy_only": text_bytes / (s.ms_ingest / 1e3) / 1e9,
                          "matches": int(n.value), "parity_ok_on_sample": ok}), flush=True)
        assert ok and nrec.value == P
        m.close()
        o.close()

    if a.list:
        # list mode: K fusion CSVs = K independent indices over the same reads (fusion_scan.rs:62-188); alternate the
        # full panel and a 40-gene subset like benchmark_res/hg38_fusion_csv_list.txt alternates cancer/druggable
        L, P = 150, a.pairs
        batch, d, d_off = device_batch(torch, synth, panel, P, L, 12, threads)
        sub = genes[:40]
        mappers = [FusionMapper.from_gene_spans(genes if k % 2 == 0 else sub) for k in range(a.list)]
        ms_sep, s, _ = run(mappers, d, d_off, P, L, a.steps)      # one gf_map_pairs_device per CSV (k_prep per CSV)
        # one gf_map_pairs_device_list call: upload / k_prep once, then per index
        K = a.list
        lib = mappers[0].lib
        out_cap = max(1 << 16, P // 4)
        d_outs = [torch.empty(out_cap * C.sizeof(gf_match), dtype=torch.uint8, device="cuda") for _ in range(K)]
        d_ns = torch.zeros(K, dtype=torch.int64, device="cuda")
        hs = (C.c_void_p * K)(*[m.m_indexer.h.value for m in mappers])
        outs = (C.c_void_p * K)(*[t.data_ptr() for t in d_outs])
        nouts = (C.c_void_p * K)(*[d_ns.data_ptr() + 8 * k for k in range(K)])
        db = gf_batch()
        db.n = P
        db.seq1, db.qual1, db.seq2, db.qual2 = (t.data_ptr() for t in d)
        db.off1 = db.off2 = d_off.data_ptr()
        db.bytes1 = db.bytes2 = P * L
        db.max_len = L
        st = torch.cuda.current_stream()

        def list_pass():
            rc = lib.gf_map_pairs_device_list(hs, K, C.byref(db), outs, out_cap, nouts, C.c_void_p(st.cuda_stream))
            assert rc == 0, lib.gf_last_error()
        for _ in range(3):
            list_pass()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(a.steps):
            list_pass()
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        n_list = [int(x) for x in d_ns.tolist()]
        # end to end from pinned host memory: gf_list_map_pairs (one upload) vs one gf_map_pairs per CSV
        from genefuserust_b200 import host as H
        import numpy as np
        hb = ReadBatch(batch.seq1, batch.qual1, batch.off1, batch.seq2, batch.qual2, batch.off2)
        got_list = H.scan_list(mappers, hb)                        # first call: grows every handle's chunk buffers
        got_sep = [m.scan_pair_end(hb) for m in mappers[:2]]
        # timed: the C ABI calls themselves (preallocated outputs), pinned host arenas
        cap_h = max(4096, P // 64)
        bufs = [(gf_match * cap_h)() for _ in range(K)]
        outs_h = (C.POINTER(gf_match) * K)(*[C.cast(b_, C.POINTER(gf_match)) for b_ in bufs])
        caps_h = (C.c_uint64 * K)(*([cap_h] * K))
        nout_h = (C.c_uint64 * K)()
        hst = hb.as_struct()
        t0 = time.perf_counter()
        rc = lib.gf_list_map_pairs(hs, K, C.byref(hst), outs_h, caps_h, nout_h)
        e2e_list = time.perf_counter() - t0
        assert rc == 0, lib.gf_last_error()
        one = C.c_uint64(0)
        t0 = time.perf_counter()
        for k in range(2):
            rc = lib.gf_map_pairs(mappers[k].m_indexer.h, C.byref(hst), bufs[k], cap_h, C.byref(one))
            assert rc == 0, lib.gf_last_error()
        e2e_sep2 = time.perf_counter() - t0
        same = all([r.astuple() for r in got_list[k]] == [r.astuple() for r in got_sep[k]] for k in range(2))
        same = same and [len(g) for g in got_list] == n_list
        sample = batch.slice(0, min(a.oracle_sample, P))
        ok = same
        for gset, k in ((genes, 0), (sub, 1 if a.list > 1 else 0)):
            o = _oracle.OracleIndex(gset)
            ok = ok and [r.astuple() for r in got_list[k] if r.pair_idx < sample.n] == o.scan(sample, threads=threads)
            o.close()
        print(json.dumps({"config": f"list mode: {a.list} fusion CSVs (alternating 136-gene / 40-gene panels) x {P} pairs 2x150, 1 B200",
                          "ms_per_list_job": ms, "pairs_per_s_whole_list": P / (ms / 1e3),
                          "pair_x_csv_per_s": P * a.list / (ms / 1e3),
                          "ms_per_list_job_one_call_per_csv": ms_sep, "pair_x_csv_per_s_one_call_per_csv": P * a.list / (ms_sep / 1e3),
                          "e2e_s_list_call (one upload, pinned host)": e2e_list, "e2e_pair_x_csv_per_s": P * a.list / e2e_list,
                          "e2e_s_per_csv_separate_calls": e2e_sep2 / 2,
                          "parity_ok_on_sample": ok,
                          "index_bytes_total": sum(int(m.m_indexer.info().device_bytes) for m in mappers)}), flush=True)
        assert ok
        for m in mappers:
            m.close()


if __name__ == "__main__":
    main()
