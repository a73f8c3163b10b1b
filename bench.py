#!/usr/bin/env python
"""bench.py -- deflate compress throughput (GB/s of input) of the B200 path, BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload text|random|zeros|pattern]

N=1 workload: BASELINE.json configs[1] -- 1 GiB order-2 Markov text, 64 KiB chunks + 32 KiB dictionary,
level 2 (dynamic Huffman).  For N>1 (torchrun, one rank per GPU) every rank holds its own 1 GiB shard of the
same stream (weak scaling; shards are independent given the 32 KiB before them, so there is no data-path
collective -- only the barrier and the max-over-ranks reduction of the timing).

One step = one pass of the whole pipeline over the rank's shard.  `value` is device-timed with CUDA events on
the library's launching stream, inputs and outputs resident in HBM; `e2e` is the same metric through the public
ZzFlateEncode call on pinned HOST buffers (H2D and D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GIB = 1 << 30
CHUNK, DICT = 65536, 32768
METRIC = "deflate_compress_input_throughput"
UNIT = "GB/s"


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms from before the input is generated; the summary uses the samples
    whose timestamps fall inside the timed region (and, if the region was shorter than the sampling jitter, the
    load-carrying warm-up right before it)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.t0 = self.t1 = self.tw = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_ready(self, timeout: float = 5.0):
        """blocks until the first sample has arrived (the timed region of the default run is under 100 ms)"""
        t = time.time()
        while self.proc is not None and not self.rows and time.time() - t < timeout:
            time.sleep(0.01)
        self.tw = time.time()                  # the warm-up starts here

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()

        def summarise(rows):
            sm, mx, reasons = [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[2])); mx.append(float(r[3]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[6:10]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
                except Exception:
                    continue
            return sm, mx, reasons

        inside = [x for x in self.rows if self.t0 is not None and self.t0 - 0.02 <= x[0] <= (self.t1 or x[0]) + 0.02]
        window = "timed region"
        if len(inside) < 3:
            lo = self.tw if self.tw is not None else (self.t0 - 1.0 if self.t0 is not None else 0.0)
            inside = [x for x in self.rows if x[0] >= lo - 0.02]
            window = "timed region and the warm-up steps before it"
        sm, mx, reasons = summarise(inside)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def make_shard(workload: str, rank: int, nbytes: int):
    """Rank's shard of the stream plus the history (dictionary + backward-extension slack) before it."""
    import numpy as np
    from zzflate_b200 import synth
    hist = 0 if rank == 0 else DICT + 288
    buf = np.empty(hist + nbytes, dtype=np.uint8)
    if workload == "text":
        segs = nbytes >> 20
        if hist:
            buf[:hist] = synth.markov_text(1 << 20, seg0=rank * segs - 1)[-hist:]
        synth.markov_text(nbytes, seg0=rank * segs, out=buf[hist:])
    else:
        full = synth.workload(workload, hist + nbytes)        # random/zeros/pattern: each rank an independent stream
        buf[:] = full
    return buf, hist


def cpu_reference_run(data, steps: int, warmup: int):
    """The reference's own CPU encoder (oracle/_ref, else the oracle port) on the host cores."""
    import oracle_lib
    cores = os.cpu_count() or 1
    n = data.size
    if oracle_lib.REF_SO.exists():
        ref = oracle_lib.reference(); kind = "reference"
        run = lambda: ref.encode(data, oracle_lib.DEFLATE, 2, threaded=True)       # hardware_concurrency() tasks
    else:
        o = oracle_lib.oracle(); kind = "port"
        run = lambda: o.stream_chunked(data, oracle_lib.DEFLATE, 2, threads=cores)[0]
    out_len = 0
    for _ in range(warmup):
        out_len = len(run())
    times = []
    for _ in range(steps):
        t = time.perf_counter(); out_len = len(run()); times.append(time.perf_counter() - t)
    return kind, cores, times, out_len


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="text", choices=["text", "random", "zeros", "pattern"])
    ap.add_argument("--level", type=int, default=2)
    ap.add_argument("--shard-mib", type=int, default=1024, help="input bytes per GPU per step (MiB)")
    ap.add_argument("--cpu-sample-mib", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the output checks that run after the timed loops")
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to zzgpu_set_option (A/B of kernel variants)")
    ap.add_argument("--format", default="zlib", choices=["zlib", "gzip", "deflate"], help="framing of the end-to-end leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lvl = {0: "level 0 stored", 1: "level 1 fixed Huffman", 2: "level 2 dynamic Huffman", 3: "level 3 dynamic Huffman"}[args.level]
    workload_name = {"text": "1 GiB order-2 Markov text per GPU (SURVEY 8d.2), 64 KiB chunks + 32 KiB dictionary, " + lvl,
                     "random": "1 GiB splitmix64 random bytes per GPU (SURVEY 8d.3), " + lvl,
                     "zeros": "1 GiB all-zero input per GPU (SURVEY 8d.4a), " + lvl,
                     "pattern": "1 GiB of a 1000-byte pattern repeated per GPU (SURVEY 8d.4b), " + lvl}[args.workload]
    config = {"workload": workload_name if args.shard_mib == 1024 else workload_name.replace("1 GiB", f"{args.shard_mib} MiB"),
              "bytes_per_gpu": args.shard_mib << 20, "chunk": CHUNK, "dict": DICT, "level": args.level,
              "sharding": f"{world} contiguous shards, no collective" if world > 1 else "single GPU",
              "l2": "inputs (>= 256 MiB per step) exceed the 126 MB L2; no explicit flush"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        from zzflate_b200 import synth
        n = min(args.cpu_sample_mib, args.shard_mib) << 20
        data = synth.workload(args.workload, n)
        kind, cores, times, out_len = cpu_reference_run(data, args.steps, args.warmup)
        t = sum(times) / len(times)
        val = n / t / 1e9
        sample = f"first {n >> 20} MiB of the workload per step, Format=Deflate level 2, threaded=true"
        emit_line({"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 3), "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config,
                          "ratio": round(out_len / n, 5),
                          "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                          "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ---------------------------------------------------------------- B200 arm
    import numpy as np
    import torch
    import zzflate_b200 as zz
    from zzflate_b200 import _lib

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    _lib.check(_lib.load().zzgpu_init(local_rank))
    for kv in args.opt:
        k, v = kv.split("=")
        _lib.check(_lib.load().zzgpu_set_option(k.encode(), int(v)))
        config.setdefault("options", {})[k] = int(v)

    nbytes = args.shard_mib << 20
    sampler = ClockSampler(local_rank)          # started early: nvidia-smi needs a moment before its first sample
    sampler.start()
    host, hist = make_shard(args.workload, rank, nbytes)
    final = rank == world - 1
    pinned_src = torch.from_numpy(host).pin_memory()
    d_src = pinned_src.cuda(non_blocking=False)
    cap = zz.bound(nbytes, args.level)
    d_dst = torch.empty(cap, dtype=torch.uint8, device="cuda")
    pinned_dst = torch.empty(cap + 64, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        return zz.deflate_device(d_src.data_ptr() + hist, nbytes, d_dst.data_ptr(), cap, level=args.level,
                                 history=hist, final=final, checksums=1)

    sampler.wait_ready()
    for _ in range(args.warmup):
        out_len, _, _, st = device_step()
    barrier()
    sampler.mark_begin()
    dev_ms, stage_ms, launches = 0.0, [0.0] * len(_lib.STAGES), 0
    stage_n = [0] * len(_lib.STAGES)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_len, _, _, st = device_step()
        dev_ms += st.device_ms
        launches += st.kernel_launches
        for i in range(len(_lib.STAGES)):
            stage_ms[i] += st.stage_ms[i]; stage_n[i] += st.stage_launches[i]
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.mark_end()
    clocks = sampler.stop()
    # ---- end to end through the public API on pinned host buffers
    e2e_ms, e2e_out = None, 0
    if not args.no_e2e:
        cfg = zz.Config({"zlib": zz.Format.Zlib, "gzip": zz.Format.Gzip, "deflate": zz.Format.Deflate}[args.format], args.level, False)
        src_ptr = pinned_src.data_ptr() + hist
        if hist:          # the public single-call API has no history argument: rank>0 encodes its shard as its own stream
            pass
        for _ in range(min(args.warmup, 2)):
            e2e_out = zz.encode_ptr(pinned_dst.data_ptr(), cap, src_ptr, nbytes, cfg)
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            e2e_out = zz.encode_ptr(pinned_dst.data_ptr(), cap, src_ptr, nbytes, cfg)
            assert e2e_out is not None
        barrier()
        e2e_ms = (time.perf_counter() - t1) * 1e3


    # ---- verification of the last timed outputs (outside the timed regions): a wrong stream must not print a number
    verified = None
    if not args.no_verify:
        import zlib
        out_len, a0, crc_v, _ = zz.deflate_device(d_src.data_ptr() + hist, nbytes, d_dst.data_ptr(), cap, level=args.level,
                                                  history=hist, final=final, checksums=3)
        got = d_dst[:out_len].cpu().numpy()
        shard = host[hist:]
        ok = {}
        d = zlib.decompressobj(-15, zdict=host[max(0, hist - DICT): hist].tobytes()) if hist else zlib.decompressobj(-15)
        pos, good = 0, True
        for lo in range(0, out_len, 64 << 20):
            piece = d.decompress(got[lo: lo + (64 << 20)].tobytes())
            good = good and piece == shard[pos: pos + len(piece)].tobytes()
            pos += len(piece)
        ok["inflate_equals_input"] = bool(good and pos == nbytes and (d.eof or not final))
        if final:
            # second, zlib-independent judge: the library's own inflater (include/decoder.h, host code)
            back = zz.ZzFlateDecode(got, zz.Format.Deflate, max_len=nbytes + 16, dictionary=host[:hist] if hist else None)
            ok["decoder_h_equals_input"] = bool(back is not None and len(back) == nbytes and np.array_equal(np.frombuffer(back, dtype=np.uint8), shard))
        ok["adler32"] = bool(a0 == zlib.adler32(shard, 0))
        ok["crc32"] = bool(crc_v == zlib.crc32(shard))
        if rank == 0:
            # rank 0's shard against the CPU restatement of the reference over the whole shard (the checker, not the product)
            import oracle_lib
            want, _ = oracle_lib.oracle().stream_chunked(shard, oracle_lib.DEFLATE, args.level, threads=os.cpu_count() or 1)
            keep = len(want) if final else len(want) - (CHUNK + 64)       # the last chunk differs when the shard is not final
            ok["rank0_bytes_equal_oracle"] = bool(out_len >= keep and got[:keep].tobytes() == want[:keep])
            ok["oracle_compared_bytes"] = int(keep)
        if e2e_ms is not None:
            e2e_bytes = pinned_dst[:e2e_out].numpy()
            wbits = {"zlib": 15, "gzip": 31, "deflate": -15}[args.format]
            d2 = zlib.decompressobj(wbits)
            pos2, good2 = 0, True
            for lo in range(0, int(e2e_out), 64 << 20):
                piece = d2.decompress(e2e_bytes[lo: lo + (64 << 20)].tobytes())
                good2 = good2 and piece == shard[pos2: pos2 + len(piece)].tobytes()
                pos2 += len(piece)
            ok["e2e_inflate_equals_input"] = bool(good2 and pos2 == nbytes and d2.eof)       # includes zlib's own Adler-32 / CRC-32 check
        meta = torch.tensor([out_len, a0, crc_v, nbytes, zlib.adler32(shard, 0), zlib.crc32(shard), int(all(v for v in ok.values() if isinstance(v, bool)))],
                            dtype=torch.int64, device="cuda")
        if dist is not None:
            metas = [torch.zeros_like(meta) for _ in range(world)]
            dist.all_gather(metas, meta)
        else:
            metas = [meta]
        metas = [m.cpu().tolist() for m in metas]
        # fold of the ranks' (checksum, length) pairs as the host driver does (zzgpu_adler32_combine / zzgpu_crc32_combine)
        # against the same fold of zlib's per-shard values through the oracle's independent combine
        if rank == 0:
            import oracle_lib
            o = oracle_lib.oracle()
            fa, fc, wa, wc = 1, 0, 1, 0
            for m in metas:
                fa = zz.combine(fa, m[1], m[3]); fc = zz.crc32_combine(fc, m[2], m[3])
                wa = o.combine(wa, m[4], m[3]); wc = o.crc32_combine(wc, m[5], m[3])
            ok["folded_adler32"] = bool(fa == wa); ok["folded_crc32"] = bool(fc == wc)
            ok["all_ranks_ok"] = bool(all(m[6] == 1 for m in metas))
            ok["stream_bytes_all_ranks"] = int(sum(m[0] for m in metas))
            verified = ok
            if not all(v for v in ok.values() if isinstance(v, bool)):
                print("VERIFICATION FAILED: " + json.dumps(ok), file=sys.stderr, flush=True)
                if dist is not None:
                    dist.destroy_process_group()
                sys.exit(3)

    vals = torch.tensor([dev_ms, wall_ms, e2e_ms or 0.0, float(launches), float(out_len)], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, wall_ms, e2e_ms_r = mx[0].item(), mx[1].item(), mx[2].item()
        launches = int(sm[3].item()); total_out = sm[4].item()
    else:
        e2e_ms_r = e2e_ms or 0.0; total_out = float(out_len)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    total_in = float(nbytes) * world
    ms_per_step = dev_ms / args.steps
    value = total_in / (ms_per_step * 1e-3) / 1e9
    ratio = total_out / total_in
    peak, peak_src = measured_peaks()
    # dominant kernel (rank 0's stage times)
    names = _lib.STAGES
    dom = max(range(len(_lib.STAGES)), key=lambda i: stage_ms[i])
    dom_avg_ms = stage_ms[dom] / max(stage_n[dom], 1)
    chunks_per_launch = ((nbytes + CHUNK - 1) // CHUNK) / max(stage_n[dom] / args.steps, 1)
    alg_bytes = chunks_per_launch * CHUNK * (1.0 + ratio)            # 1 read + r written per input byte
    achieved = alg_bytes / (dom_avg_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of that kernel from the committed ncu --set full capture (not
    # measurable inside a plain run); the source of the number travels with it
    traffic, traffic_src = None, None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists() and args.workload == "text" and args.level == 2 and args.shard_mib == 1024:
        try:
            tj = json.loads(tp.read_text())
            traffic = tj.get(names[dom]); traffic_src = "profiles/traffic.json: " + str(tj.get("_source"))
        except Exception:
            traffic = None
    line = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config, "ratio": round(ratio, 5),
            "wall_ms_per_step": round(wall_ms / args.steps, 3),
            "stage_ms_per_step": {names[i]: round(stage_ms[i] / args.steps, 3) for i in range(len(_lib.STAGES)) if stage_n[i]},

            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 5), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": int(alg_bytes), "avg_launch_ms": round(dom_avg_ms, 4),
                         "whole_path_frac": round(value * (1.0 + ratio) / peak, 5)},
            "gpu_launches": launches, "clocks": clocks, "verified": verified}
    if e2e_ms is not None:
        e2e_val = total_in / (e2e_ms_r / args.steps * 1e-3) / 1e9
        line["e2e"] = {"value": round(e2e_val, 3), "unit": UNIT, "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(e2e_out or 0),
                       "ms_per_step": round(e2e_ms_r / args.steps, 3), "api": f"ZzFlateEncode(Format={args.format}, level, threaded=false) on pinned host buffers (checksum of the trailer computed on the GPU inside the call)"}
    if not args.no_cpu_baseline:
        n_s = min(args.cpu_sample_mib << 20, nbytes)
        kind, cores, times, cpu_out = cpu_reference_run(host[hist: hist + n_s], 3, 1)
        best = min(times)
        line["cpu_baseline"] = {"value": round(n_s / best / 1e9, 4), "unit": UNIT, "cores": cores, "kind": kind,
                                "sample": f"first {n_s >> 20} MiB of rank 0's shard, ZzFlateEncode Format=Deflate level 2 threaded=true, best of 3",
                                "ratio": round(cpu_out / n_s, 5)}
    emit_line(line)
    if dist is not None:
        dist.destroy_process_group()


class _StdoutToStderr:
    """stdout carries exactly one JSON line: while the bench runs, file descriptor 1 points at stderr (NCCL and the
    libraries under torch print version banners to it), and the real stdout comes back for the final print."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1); os.close(self.saved); self.saved = None

    def __exit__(self, *exc):
        self.restore()
        return False


_GUARD = None


def emit_line(obj) -> None:
    if _GUARD is not None:
        _GUARD.restore()
    print(json.dumps(obj), flush=True)


if __name__ == "__main__":
    with _StdoutToStderr() as _GUARD:
        main()
