#!/usr/bin/env python3
"""bench.py — read pairs/s of the fusion-matching hot path on synthetic 2x150 bp pairs vs the cancer.csv-shaped
panel (BASELINE.json configs[1]), device-timed, with the HBM roofline of the dominant kernel, the end-to-end
number through the C ABI from pinned host buffers, and the CPU oracle timed on the box's own cores.

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (C++ restatement: no
                                                           # Rust toolchain in this image)
One rank per GPU under torchrun (weak scaling: every rank maps its own --pairs pairs, index replicated, no
data-path collective).  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "read_pairs_per_sec_matched"
UNIT = "pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="pairs per GPU per step (configs[1]: 10M)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--seed", type=int, default=12)
    ap.add_argument("--panel-scale", type=float, default=1.0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU-baseline sample (0 = auto, ~15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic 2x{a.read_len}bp, {a.pairs} pairs/GPU vs cancer.csv-shaped panel "
            f"(136 genes, {15.1 * a.panel_scale:.1f} Mbases) on synthetic contigs")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None, t_end=None):
        """clocks of the samples that arrived inside [t_begin, t_end] (host clock); when the timed region is shorter than
        the sampling period the samples of the whole loaded phase (warm-up .. extra steps, the same kernels) are used and
        `window` says so"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(lines):
            sm, mx, reasons = [], [], set()
            for _t, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            sm.sort()
            return sm, mx, sorted(reasons)
        window = "timed region"
        inside = [x for x in self.lines if t_begin is None or (t_begin - 0.01 <= x[0] <= t_end + 0.03)]
        sm, mx, reasons = summarise(inside)
        if len(sm) < 2:
            window = "warm-up + timed region + further identical steps (timed region shorter than the sampling period)"
            sm, mx, reasons = summarise(self.lines)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons, "window": window}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any"""
    p = os.path.join(ROOT, "profiles", "screen_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def bind_to_gpu_numa_node(torch, local_rank):
    """Several ranks on one box: run this rank (and therefore first-touch / pin its host buffers) on the CPUs of the NUMA
    node its GPU hangs off, so that the H2D streams of the ranks do not all cross the socket interconnect.  Best effort:
    returns the node, or None when the topology cannot be read (GF_BENCH_NUMA=0 disables it)."""
    if os.environ.get("GF_BENCH_NUMA", "1") == "0":
        return None
    try:
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_cpu_oracle(panel, batch, threads, sample_pairs):
    """times the CPU oracle (test infrastructure; here ONLY as the reported CPU baseline) on a bounded sample"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle
    sample = batch.slice(0, min(sample_pairs, batch.n))
    t0 = time.perf_counter()
    oidx = _oracle.OracleIndex(panel.genes())
    t_index = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = oidx.scan(sample, threads=threads)
    dt = time.perf_counter() - t0
    oidx.close()
    return sample.n / dt, dt, len(res), t_index


def _only_json_on_stdout():
    """Libraries (NCCL's version banner, nvcc) write to the C-level stdout; the contract is ONE JSON line there.
    Everything else goes to stderr; the JSON line is written to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    a = parse_args()
    json_out = _only_json_on_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import numpy as np
    import __graft_entry__ as ge
    from genefuserust_b200 import synth
    from genefuserust_b200.batch import ReadBatch

    # ------------------------------------------------------------------ reference arm: CPU only, rank 0 only
    if a.impl == "reference":
        if rank != 0:
            return
        ge.build()
        panel = synth.make_panel(scale=a.panel_scale)
        threads = cpu_threads()
        n_sample = a.cpu_sample or 200_000
        batch = synth.generate_pairs(panel, n_sample, read_len=a.read_len, seed=a.seed, threads=threads)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import _oracle
        oidx = _oracle.OracleIndex(panel.genes())
        for _ in range(a.warmup):
            oidx.scan(batch.slice(0, min(20000, batch.n)), threads=threads)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            oidx.scan(batch, threads=threads)
        dt = time.perf_counter() - t0
        val = batch.n * a.steps / dt
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int64", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": f"each step = the first {batch.n} pairs of the workload"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{batch.n} pairs per step x {a.steps} steps, C++ restatement of the Rust "
                                       "CPU path (oracle/), packs of 1000 pairs over all host threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
        return

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ge.build()
    from genefuserust_b200._abi import gf_batch, gf_map_stats, gf_match
    from genefuserust_b200.host import FusionMapper

    if numa is not None:   # the affinity now covers one NUMA node: share it among the ranks bound to that node
        import glob
        n_nodes = max(1, len(glob.glob("/sys/devices/system/node/node[0-9]*")))
        threads = max(1, cpu_threads() // max(1, -(-world // n_nodes)))
    else:
        threads = max(1, cpu_threads() // max(1, world))
    panel = synth.make_panel(scale=a.panel_scale)
    # the first create also pays CUDA context + module load; time a second one for the steady-state index build
    genes = panel.genes()
    t0 = time.perf_counter()
    FusionMapper.from_gene_spans(genes, device=local_rank).close()
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    mapper = FusionMapper.from_gene_spans(genes, device=local_rank)
    t_index = time.perf_counter() - t0
    info = mapper.m_indexer.info()
    lib, h = mapper.lib, mapper.m_indexer.h

    # this rank's shard of the counter-based workload, generated straight into pinned host memory.  Shards above 20 M
    # pairs (BASELINE config 3: 100 M pairs = 60 GB with qualities) are generated and uploaded in 10 M-pair pieces through
    # one reusable pinned buffer; the e2e / CPU legs then use the first piece.
    P, L = a.pairs, a.read_len
    dev = torch.device("cuda", local_rank)
    HP = P if P <= 20_000_000 else 10_000_000          # pairs held in pinned host memory
    pinned = [torch.empty(HP * L, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
    t0 = time.perf_counter()
    if HP == P:
        batch = synth.generate_pairs(panel, P, read_len=L, seed=a.seed, first=rank * P, threads=threads,
                                     out=tuple(t.numpy() for t in pinned))
        d_arr = [t.to(dev, non_blocking=True) for t in pinned]
    else:
        d_arr = [torch.empty(P * L, dtype=torch.uint8, device=dev) for _ in range(4)]
        for lo in range(P - P % HP if P % HP else P - HP, -1, -HP):     # descending, so the first piece stays in `pinned`
            cn = min(HP, P - lo)
            batch = synth.generate_pairs(panel, cn, read_len=L, seed=a.seed, first=rank * P + lo, threads=threads,
                                         out=tuple(t.numpy()[:cn * L] for t in pinned))
            for dt_, pt in zip(d_arr, pinned):
                dt_[lo * L:(lo + cn) * L].copy_(pt[:cn * L], non_blocking=True)
            torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    off_host = torch.from_numpy(batch.off1.view(np.int64)).pin_memory()
    batch.off1 = off_host.numpy().view(np.uint64)
    batch.off2 = batch.off1
    d_off = (torch.arange(P + 1, dtype=torch.int64, device=dev) * L) if HP != P else off_host.to(dev, non_blocking=True)
    out_cap = max(1 << 16, P // 4)
    d_out = torch.empty(out_cap * C.sizeof(gf_match), dtype=torch.uint8, device=dev)
    d_nout = torch.zeros(1, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()

    db = gf_batch()
    db.n = P
    db.seq1, db.qual1, db.seq2, db.qual2 = (t.data_ptr() for t in d_arr)
    db.off1 = db.off2 = d_off.data_ptr()
    db.bytes1 = db.bytes2 = P * L
    db.max_len = L
    stream = torch.cuda.current_stream()

    def step_device():
        rc = lib.gf_map_pairs_device(h, C.byref(db), d_out.data_ptr(), out_cap, d_nout.data_ptr(),
                                     C.c_void_p(stream.cuda_stream))
        if rc != 0:
            raise RuntimeError(lib.gf_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()          # before the warm-up: nvidia-smi needs ~100 ms to deliver its first line
    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    screen_ms = []
    exact_ms = []
    parts_ms = []
    barrier()
    t_begin = time.perf_counter()
    ev0.record(stream)
    for _ in range(a.steps):
        step_device()
    ev1.record(stream)
    barrier()
    t_end = time.perf_counter()
    ms_total = ev0.elapsed_time(ev1)
    # per-kernel duration of the last step from the library's own events on the launching stream
    st = gf_map_stats()
    lib.gf_get_map_stats(h, C.byref(st))
    screen_ms.append(st.ms_screen)
    exact_ms.append(st.ms_exact)
    parts_ms.append((st.ms_prep, st.ms_seed, st.ms_diag, st.ms_scan))
    # a few extra single steps to average the dominant kernel's launch duration live
    for _ in range(min(3, a.steps)):
        step_device()
        s2 = gf_map_stats()
        lib.gf_get_map_stats(h, C.byref(s2))
        screen_ms.append(s2.ms_screen)
        exact_ms.append(s2.ms_exact)
        parts_ms.append((s2.ms_prep, s2.ms_seed, s2.ms_diag, s2.ms_scan))
    # a short timed region (small --steps) can end before nvidia-smi has delivered a line: keep the same kernels running
    # until a few samples exist (outside the timed numbers; `clocks.window` says which samples were used)
    t_extra = time.perf_counter()
    while len(sampler.lines) < 4 and time.perf_counter() - t_extra < 1.0:
        step_device()
        torch.cuda.synchronize()
    clocks = sampler.stop(t_begin, t_end)
    n_matches = int(d_nout.item())

    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(n_matches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total_max = float(t.item())
    value = world * P * a.steps / (ms_total_max / 1000.0)

    # ---- end to end through the public C-ABI call with HOST (pinned) buffers: H2D + kernels + D2H inside
    e2e = None
    if not a.no_e2e:
        hb = batch.as_struct()
        out_host = (gf_match * out_cap)()
        n_out = C.c_uint64(0)

        def step_host():
            rc = lib.gf_map_pairs(h, C.byref(hb), out_host, out_cap, C.byref(n_out))
            if rc != 0:
                raise RuntimeError(lib.gf_last_error().decode())
        step_host()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(a.steps, 3))
        for _ in range(e2e_steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        assert HP != P or int(n_out.value) == n_matches, (n_out.value, n_matches)
        hs = gf_map_stats()
        lib.gf_get_map_stats(h, C.byref(hs))
        e2e = {"value": world * batch.n * e2e_steps / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(hs.h2d_bytes), "d2h_bytes_per_step": int(hs.d2h_bytes),
               "zero_copy_qualities": bool(hs.zero_copy_qual),
               "host_buffer_bytes_per_step": 4 * batch.n * L + 2 * 8 * (batch.n + 1),
               "pairs_per_step": int(batch.n),
               "steps": e2e_steps, "timing": "host wall clock around the synchronous C-ABI call, max over ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_screen): algorithmic bytes / measured launch duration
    peak, peak_src = measured_peak()
    # bases of every mapped sequence + one 32-byte sector per pass-1 probe (qualities are only touched
    # where fast_merge's decision depends on them, so they are not counted)
    alg_bytes = st.seq_bytes + 32 * st.n_probes_pass1
    k_ms = sum(screen_ms) / len(screen_ms)
    achieved = alg_bytes / (k_ms / 1000.0) / 1e9
    traffic = ncu_traffic()
    names = ("k_prep", "k_seed", "k_diag", "k_scan")
    part = [sum(p[i] for p in parts_ms) / len(parts_ms) for i in range(4)]
    per_kernel = {}
    for i, nm in enumerate(names):
        e = {"ms": part[i]}
        if traffic and nm in traffic.get("kernels", {}):
            e.update(traffic["kernels"][nm])           # ncu --set full, profiles/: dram bytes, issue / L1TEX utilisation
            if part[i] > 0 and "dram_bytes" in e:
                e["dram_gbs_live"] = e["dram_bytes"] / (part[i] / 1000.0) / 1e9
        per_kernel[nm] = e
    dram = traffic["dram_bytes_per_launch"] if traffic else None
    roofline = {"bound": "hbm",
                "kernel": "screen = k_prep + k_seed + k_diag + k_scan (fast_merge + conservative pass 1 of Indexer::map_read for "
                          "every pair; 4 launches per step, timed together with CUDA events on the launching stream)",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "traffic": dram,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_ms,
                "dram_frac": (dram / (k_ms / 1000.0) / 1e9 / peak) if dram else None,
                "per_kernel": per_kernel,
                "note": "achieved = SURVEY 8(d) algorithmic bytes (bases of every mapped sequence + one 32-byte HBM sector per "
                        "pass-1 probe) / measured time.  The probes are answered from an L2-resident filter and 2-bit gene "
                        "planes instead of an HBM hash table, so the screen moves far fewer DRAM bytes (traffic, dram_frac) "
                        "than the model charges and frac can exceed 1: the model is the common yardstick, not the limiter.  "
                        "What limits the kernels now is instruction issue (k_prep) and L1TEX gather throughput (k_scan, "
                        "k_diag), see per_kernel and profiles/.",
                "kernel_share_of_step": k_ms / (ms_total_max / a.steps),
                "exact_verify_ms": sum(exact_ms) / len(exact_ms)}

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        n_s = a.cpu_sample
        if not n_s:
            r, dt, _, _ = run_cpu_oracle(panel, batch, cpu_threads(), 20000)
            n_s = int(max(20000, min(P, r * 15)))
        r, dt, nm, t_oidx = run_cpu_oracle(panel, batch, cpu_threads(), n_s)
        cpu = {"value": r, "unit": UNIT, "cores": cpu_threads(), "kind": "port",
               "sample": f"first {n_s} pairs of the same workload, {dt:.1f} s; C++ restatement of the Rust CPU path "
                         f"(no Rust toolchain here), packs of 1000 pairs over all host threads; index build {t_oidx:.1f} s "
                         "not included"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_total_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int64", "data": "synthetic",
        "config": {"workload": workload_name(a), "pairs_per_gpu": P, "read_len": L, "seed": a.seed,
                   "index": {"keys": info.n_keys, "sites": info.n_sites, "table_bytes": info.table_bytes,
                             "build_ms": info.build_ms},
                   "l2": f"inputs ({4 * P * L / 1e9:.0f} GB/GPU) and table (0.5 GB) both exceed the 126 MB L2; no flush needed",
                   "sharding": "pairs sharded by rank, index replicated, no data-path collective"},
        "matches_per_step": float(cnt.item()),
        "survivors_per_step": int(st.n_survivors),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(st.kernel_launches) * a.steps,
        "roofline": roofline, "cpu_baseline": cpu,
        "setup": {"numa_node": numa, "index_create_s": t_index, "first_index_create_s_incl_cuda_init": t_first, "generate_s": t_gen,
                  "host_threads": threads},
    }
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
