#!/usr/bin/env python3
"""bench.py -- the reference's VAD -> MFCC -> DTW hot path (spch_recg, Src/APP/main.c:249-296) on B200.

One "step" = one pass of the whole path (noise_atap + VAD + get_mfcc + dtw x T + argmin) over one batch of
synthetic utterances (BASELINE.json configs[1]: 65 536 x 1 s utterances, 12 MFCC, 20 templates, per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W]          one JSON line (rank 0)
  python bench.py --impl reference ...                         the reference's own C on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...            one rank per GPU, utterances sharded (weak
                                                               scaling), one NCCL all-gather of the scores

`value`  : utterances/s, whole job, inputs resident in HBM (device-pointer C-ABI, CUDA events, max over ranks)
`e2e`    : the same metric through the host-buffer C-ABI call sr_recognise_batch (pinned host PCM in,
           command index + match distance + status out; H2D and D2H inside the timed region)
`roofline`: the dominant kernel (mfcc_kernel), algorithmic bytes (2*U_seg + 24*F + 4 per
           utterance) / its event-timed duration inside the timed steps, vs the measured HBM peak
`cpu_baseline`: oracle/_ref/libref.so (the reference's own C, one process per core) -- or the oracle port --
           on a bounded sample of the same batch, timed on this box's host cores; the sample doubles as a
           bit-exact parity check of the GPU results.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "stm32-speech-recognition_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

U = 8000                 # 1 s @ 8 kHz
N_LEN = 2400             # 300 ms noise window (ADC.H:10-11)
SEED = 0x5EED0000
TPL_SEED = 0x7E3A0000
TAGS = {0: "vad", 1: "mfcc", 2: "status", 3: "best_init", 4: "dtw", 5: "best_final", 6: "dtw_band"}


def usable_cores():
    """host threads this process may actually use: affinity mask capped by the cgroup CPU quota"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = max(1, min(n, int(float(q) / float(p) + 0.5)))
    except Exception:
        pass
    return n


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled every 10 ms through NVML while the timed region runs (the region is often
    only ~0.1 s long, too short for `nvidia-smi -lms`); falls back to nvidia-smi when pynvml is unusable"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu, self.t_mark, self.t_end = [], None, gpu_index, None, None
        self.nv, self.stop_flag, self.t = None, False, None

    def mark(self):
        """start of the timed region: samples before it (warm-up, also under load) are used only if the
        region itself was too short to be sampled"""
        self.t_mark = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    idx = self.gpu
            self.nv = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nv
        R = nv
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = float("nan")
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = float("nan")
                flag = lambda m: "Active" if rs & m else "Not Active"
                self.rows.append([time.time(), "%g" % sm, "%g" % mx, "%g" % pw, flag(R.nvmlClocksEventReasonHwSlowdown),
                                  flag(R.nvmlClocksEventReasonHwThermalSlowdown), flag(R.nvmlClocksEventReasonSwThermalSlowdown),
                                  flag(R.nvmlClocksEventReasonSwPowerCap)])
            except Exception:
                pass
            time.sleep(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def stop(self):
        if self.nv is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        if self.nv is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            how = "NVML, 10 ms period"
        else:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            how = "nvidia-smi -lms 100"
        rows = [r for r in self.rows if len(r) >= 8]
        t1 = self.t_end if self.t_end is not None else float("inf")
        timed = [r for r in rows if self.t_mark is not None and self.t_mark <= r[0] <= t1]
        window = "timed region"
        if len(timed) < 3:                               # region shorter than a few samples
            timed, window = rows[-max(3, len(timed)):], "warm-up + timed region (timed region too short to sample)"
        num = lambda x: x.replace(".", "").isdigit()
        sm = [float(r[1]) for r in timed if num(r[1])]
        mx = [float(r[2]) for r in timed if num(r[2])]
        pw = [float(r[3]) for r in timed if num(r[3])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[4 + i] == "Active" for r in timed)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window, "how": how, "reasons": reasons}


# ---------------------------------------------------------------------------------------------------
def cpu_reference_time(pcm_sample, bank, T, cores, want_outputs=True):
    """Run the reference's C (libref.so, one PROCESS per core: it keeps statics) or the oracle port (threads)
    over pcm_sample; returns (seconds = slowest worker, kind, outputs or None)."""
    import oracle_bind as ob
    S = pcm_sample.shape[0]
    if ob.have_ref():
        import multiprocessing as mp
        ctx = mp.get_context("fork")
        bounds = [(S * k // cores, S * (k + 1) // cores) for k in range(cores)]
        with ctx.Pool(cores, initializer=_ref_init, initargs=(pcm_sample, bank, T)) as pool:
            pool.map(_ref_warm, range(cores))
            res = pool.map(_ref_work, bounds)
        secs = max(r[0] for r in res)
        out = None
        if want_outputs:
            out = {k: np.concatenate([r[1][k] for r in res]) for k in res[0][1]}
        return secs, "reference", out
    o = ob.port()
    o.recognise_batch(pcm_sample[: min(S, 64)], N_LEN, bank, T, 4096, nthreads=cores)
    t0 = time.perf_counter()
    out = o.recognise_batch(pcm_sample, N_LEN, bank, T, 4096, nthreads=cores)
    return time.perf_counter() - t0, "port", out


_G = {}


def _ref_init(pcm, bank, T):
    import oracle_bind as ob
    _G["o"], _G["pcm"], _G["bank"], _G["T"] = ob.ref(), pcm, bank, T


def _ref_warm(_):
    _G["o"].recognise_batch(_G["pcm"][:16], N_LEN, _G["bank"], _G["T"], 4096)
    return 0


def _ref_work(b):
    lo, hi = b
    x = np.ascontiguousarray(_G["pcm"][lo:hi])
    t0 = time.perf_counter()
    out = _G["o"].recognise_batch(x, N_LEN, _G["bank"], _G["T"], 4096)
    dt = time.perf_counter() - t0
    keep = {k: out[k] for k in ("seg_off", "best_idx", "best_dis", "cmd", "status", "score")}
    keep["frames"] = out["ftr"]["frm_num"].astype(np.uint32)
    return dt, keep


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation on this box's host cores (one process per
    core, persistent pool), each step a bounded sample of the configs[1] batch"""
    if rank != 0:
        return
    import sr_b200
    import oracle_bind as ob
    cores = args.ref_procs if args.ref_procs > 0 else usable_cores()
    T = args.templates
    S = min(args.batch, max(args.ref_sample_per_core, 1) * cores)
    pcm = sr_b200.synth_pcm_host(S, U, SEED)
    tpl = sr_b200.synth_pcm_host(T, U, TPL_SEED)
    e = ob.best_oracle().recognise_batch(tpl, N_LEN, None, 0, 4096)
    bank = sr_b200.make_bank(e["ftr"])
    secs, frames = [], 0
    if ob.have_ref():
        import multiprocessing as mp
        kind = "reference"
        bounds = [(S * k // cores, S * (k + 1) // cores) for k in range(cores)]
        with mp.get_context("fork").Pool(cores, initializer=_ref_init, initargs=(pcm, bank, T)) as pool:
            for i in range(args.warmup + args.steps):
                res = pool.map(_ref_work, bounds, chunksize=1)
                if i >= args.warmup:
                    secs.append(max(r[0] for r in res))
                    frames = int(sum(int(r[1]["frames"].sum()) for r in res))
    else:
        kind = "port"
        o = ob.port()
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            out = o.recognise_batch(pcm, N_LEN, bank, T, 4096, nthreads=cores)
            if i >= args.warmup:
                secs.append(time.perf_counter() - t0)
                frames = int(out["ftr"]["frm_num"].sum())
    ms = 1e3 * float(np.mean(secs))
    val = S / (ms / 1e3)
    sample = "first %d of %d utterances per step, %d worker %s, step time = slowest worker" % (
        S, args.batch, cores, "processes" if kind == "reference" else "threads")
    print(json.dumps({
        "impl": "reference", "metric": "utterances/s", "value": val, "unit": "utterances/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 fixed-point (s16 FFT, u32 energies)", "data": "synthetic",
        "config": {"workload": "configs[1]: 65536 x 1 s utterances (8 kHz u16), 12 MFCC, %d templates; "
                               "the CPU arm times a bounded sample per step" % T,
                   "utterances_per_step": S, "samples_per_utterance": U, "templates": T},
        "mfcc_frames_per_s": frames / (ms / 1e3),
        "cpu_baseline": {"value": val, "unit": "utterances/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "utterances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_kernel_workload(args):
    """--workload mfcc | dtw | dtw_band: the single-kernel configurations of BASELINE.json (configs[1] "MFCC kernel
    roofline" with the fixed segment of SURVEY 8d, configs[2] DTW 65 536 x 200). One GPU, one JSON line each.
    These are secondary lines for the BASELINE.md table; the driver's contract line is the default workload."""
    import torch
    import sr_b200
    import oracle_bind as ob
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)
    h = sr_b200.Handle(0)
    h.set_stream(stream.cuda_stream)
    h.set_dtw_variant(args.dtw_variant)
    B = args.batch
    cores = args.ref_procs if args.ref_procs > 0 else usable_cores()
    peak, peak_src = peaks()
    if args.workload == "stream":
        # BASELINE configs[4] on this one GPU (the default bench line carries the sharded form as `config4_stream`)
        T = args.templates
        tpl = torch.empty((T, U), dtype=torch.int16, device=dev)
        sr_b200.synth_pcm_dev(tpl.data_ptr(), T, U, TPL_SEED, 1, stream.cuda_stream)
        tftr = torch.zeros((T, 2860), dtype=torch.uint8, device=dev)
        h.set_bank_dev(0, 0, 4096)
        h.recognise_dev(tpl.data_ptr(), U, T, N_LEN, ftr=tftr.data_ptr())
        bank = torch.full((T, 4096), 255, dtype=torch.uint8, device=dev)
        bank[:, :2860] = tftr
        bank[:, 0], bank[:, 1] = 12345 & 0xFF, 12345 >> 8
        torch.cuda.synchronize(dev)
        line = run_stream_shard(args, torch, None, sr_b200, h, dev, stream, 0, 1, 0, bank, T)
        line.update({"data": "synthetic", "config": {"workload": line.pop("workload")}, "gpu_launches": h.launch_count()})
        print(json.dumps(line))
        return
    if args.workload in ("mfcc", "mfcc_b"):
        geom_b = args.workload == "mfcc_b"               # GEOM_B extension (200/80/256): parity unpinned, own oracle
        if geom_b:
            h.set_geometry(1)
        pcm = torch.empty((B, U), dtype=torch.int16, device=dev)
        sr_b200.synth_pcm_dev(pcm.data_ptr(), B, U, SEED, 1, stream.cuda_stream)
        seg = torch.tensor([80, 8000], dtype=torch.int32, device=dev).repeat(B, 1).contiguous()
        atap_h = np.zeros(B, sr_b200.ATAP_DTYPE)
        atap_h["mid_val"] = 2048
        atap = torch.from_numpy(atap_h.view(np.uint8).reshape(B, 12)).to(dev)
        ftr = torch.zeros((B, 2860), dtype=torch.uint8, device=dev)
        run = lambda: h.mfcc_dev(pcm.data_ptr(), U, B, seg.data_ptr(), 2, atap.data_ptr(), ftr.data_ptr())
        nfr = (7920 - 200) // 80 + 1 if geom_b else 98
        units, unit_name = B * float(nfr), "MFCC frames/s"
        bytes_per_launch = B * (2.0 * 7921 + 24 * nfr + 4)
        S = min(B, 64 * cores)
        o = ob.best_oracle()
        pcm_s = pcm[:S].cpu().numpy().view(np.uint16)
        seg_s = np.tile(np.array([80, 8000], np.uint32), (S, 1))
        t0 = time.perf_counter()
        if geom_b:
            o = ob.port()
            ref = o.mfcc_geom_b_batch(pcm_s, seg_s, atap_h[:S])
        else:
            ref = o.mfcc_batch(pcm_s, seg_s, atap_h[:S], nthreads=cores) if o.name == "oracle-port" else o.mfcc_batch(pcm_s, seg_s, atap_h[:S])
        cpu_s = time.perf_counter() - t0
        cpu = {"value": S * nfr / cpu_s, "unit": unit_name, "cores": cores if o.name == "oracle-port" else 1, "kind": "port" if o.name == "oracle-port" else "reference",
               "sample": "first %d utterances (fixed segment), single call" % S}
        check = lambda: ob.ftr_equal(ftr[:S].cpu().numpy().view(sr_b200.FTR_DTYPE).reshape(-1), ref)
        cfg = "configs[1] kernel view: 65536 x 1 s, fixed segment [80,8000) -> 98 frames/utt, mid 2048 (SURVEY 8d config 2)"
        if geom_b:
            cpu["cores"], cpu["kind"] = 1, "port"
            checker = "own restatement sro_mfcc_geom_b (GEOM_B is an extension: parity unpinned by the reference)"
            cfg = "GEOM_B extension (200/80/256, BASELINE configs[0] framing): 65536 x 1 s, fixed segment [80,8000) -> %d frames/utt" % nfr
    else:
        T = args.templates if args.templates != 20 else 200
        fin = torch.from_numpy(sr_b200.synth_ftr_host(B, 0xD7A00000, 50, 100)).to(dev)
        bank_h = sr_b200.synth_ftr_host(T, 0xD7A10000, 50, 100, stride=4096)
        bank = torch.from_numpy(bank_h).to(dev)
        h.set_bank_dev(bank.data_ptr(), T, 4096)
        score = torch.zeros((B, T), dtype=torch.int32, device=dev)
        bidx = torch.zeros(B, dtype=torch.int32, device=dev)
        bdis = torch.zeros(B, dtype=torch.int32, device=dev)
        band = args.workload == "dtw_band"
        flags = sr_b200.DTW_BAND if band else 0
        run = lambda: h.dtw_dev(fin.data_ptr(), B, flags, 10, score.data_ptr(), bidx.data_ptr(), bdis.data_ptr())
        S = min(B, 4 * cores)
        o = ob.port()
        fin_s = fin[:S].cpu().numpy().view(sr_b200.FTR_DTYPE).reshape(-1)
        t0 = time.perf_counter()
        ref, cells = o.dtw_batch(fin_s, bank_h, T, 4096, band_r=10 if band else -1, nthreads=cores)
        cpu_s = time.perf_counter() - t0
        checker = "oracle port (the reference has no banded DP: parity unpinned)" if band else "oracle port"
        if not band and ob.have_ref():                   # greedy walk: parity against the reference's own dtw (libref.so)
            ref, _ = ob.ref().dtw_batch(fin_s, bank_h, T, 4096)
            checker = "reference C (oracle/_ref/libref.so: DTW.C compiled unmodified)"
        cells_per_pair = cells / float(S * T)
        units, unit_name = B * T * cells_per_pair, "DTW cells/s (%s)" % ("lattice points in the r=10 band" if band else "get_dis evaluations of the greedy walk")
        bytes_per_launch = B * (24.0 * 75 + 4 + 4 * T) + T * (4 + 24.0 * 75)
        cpu = {"value": cells / cpu_s, "unit": unit_name, "cores": cores, "kind": "port",
               "sample": "first %d utterances x %d templates, %d threads" % (S, T, cores)}
        check = lambda: bool(np.array_equal(score[:S].cpu().numpy().view(np.uint32), ref))
        cfg = "configs[2]: 65536 utterances x %d templates, 50..100 frames each, %s" % (T, "Sakoe-Chiba DP r=10 (extension, parity unpinned)" if band else "reference greedy walk")
    for _ in range(max(args.warmup, 3)):
        run()
    torch.cuda.synchronize(dev)
    h.timing_enable(4 * args.steps + 8)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = h.launch_count()
    ev0.record(stream)
    for _ in range(args.steps):
        run()
    ev1.record(stream)
    torch.cuda.synchronize(dev)
    ms = ev0.elapsed_time(ev1) / args.steps
    recs = h.timing_collect()
    kms = float(np.mean([m for t, m in recs if t in (1, 4, 6)])) if recs else ms
    ach = bytes_per_launch / (kms * 1e-3) / 1e9
    print(json.dumps({
        "metric": unit_name, "value": units / (ms * 1e-3), "unit": unit_name, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32 fixed-point", "data": "synthetic", "config": {"workload": cfg, "batch": B},
        "utterances_per_s": B / (ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "peak_source": peak_src, "kernel_ms": kms, "algorithmic_bytes_per_launch": bytes_per_launch},
        "cpu_baseline": cpu, "parity_vs_cpu_sample": check(), "parity_checker": locals().get("checker", cpu.get("kind")),
        "gpu_launches": h.launch_count() - l0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="utterances per GPU per step")
    ap.add_argument("--templates", type=int, default=20)
    ap.add_argument("--cpu-sample-per-core", type=int, default=4096, help="cpu_baseline leg of the B200 arm (one shot)")
    ap.add_argument("--ref-sample-per-core", type=int, default=128, help="--impl reference: utterances per core per step")
    ap.add_argument("--ref-procs", type=int, default=0, help="CPU worker count (0 = usable cores: affinity capped by cgroup quota)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="recognise", choices=["recognise", "mfcc", "mfcc_b", "dtw", "dtw_band", "stream"])
    ap.add_argument("--streams", type=int, default=8192)
    ap.add_argument("--dtw-variant", type=int, default=-1, help="greedy dtw kernel: 0 static, 1 dynamic pair scheduling, -1 library default")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3], help="1 = BASELINE configs[1] per GPU (default); 3 = configs[3]: 131072 utterances per GPU x 50 templates")
    ap.add_argument("--no-config3", action="store_true", help="multi-GPU runs: skip the appended configs[3] pass")
    ap.add_argument("--no-stream", action="store_true", help="skip the appended configs[4] streaming pass")
    ap.add_argument("--samples", type=int, default=8000, help="samples per utterance (8000 = BASELINE's 1 s; 16000 = the reference's native 2 s buffer)")
    args = ap.parse_args()
    global U
    U = args.samples

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload != "recognise":
        if rank == 0:
            run_kernel_workload(args)
        return

    import sr_b200
    # NUMA: run this rank (and every thread it creates: CUDA's, the packer pool's) on the socket its GPU hangs off, before
    # anything allocates; pinned buffers below come from sr_host_alloc_dev (pages on that node)
    orig_affinity = os.sched_getaffinity(0)
    bound_node = sr_b200.lib().sr_bind_thread_to_device(local) if sr_b200.lib().sr_device_count() > local else -1
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.config == 3:                              # BASELINE configs[3]: 1 048 576 utterances / 8 GPUs, 50 templates
        args.batch, args.templates = 131072, 50
    B, T = args.batch, args.templates
    stream = torch.cuda.Stream(dev)                 # every kernel, copy and event of the bench runs here
    torch.cuda.set_stream(stream)
    h = sr_b200.Handle(local)
    h.set_stream(stream.cuda_stream)
    h.set_dtw_variant(args.dtw_variant)
    if world > 1:                                   # the exchange step lives behind the C-ABI (sr_comm_*): NCCL id via torch
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(sr_b200.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        h.comm_create(rank, world, bytes(idt.cpu().numpy().tobytes()))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def make_workload(Bw, Tw):
        """synthetic inputs on the device (byte-identical to the host generator), enrolled bank, output buffers"""
        w = {"B": Bw, "T": Tw}
        w["pcm"] = torch.empty((Bw, U), dtype=torch.int16, device=dev)
        sr_b200.synth_pcm_dev(w["pcm"].data_ptr(), Bw, U, SEED + rank * Bw, 1, stream.cuda_stream)
        tpl = torch.empty((Tw, U), dtype=torch.int16, device=dev)
        sr_b200.synth_pcm_dev(tpl.data_ptr(), Tw, U, TPL_SEED, 1, stream.cuda_stream)
        # enrolment (save_mdl, main.c:121-138): template features -> flash-layout bank
        tftr = torch.zeros((Tw, 2860), dtype=torch.uint8, device=dev)
        h.set_bank_dev(0, 0, 4096)
        h.recognise_dev(tpl.data_ptr(), U, Tw, N_LEN, ftr=tftr.data_ptr())
        bank = torch.full((Tw, 4096), 255, dtype=torch.uint8, device=dev)
        bank[:, :2860] = tftr
        bank[:, 0], bank[:, 1] = 12345 & 0xFF, 12345 >> 8
        w["bank"] = bank
        w["seg"] = torch.zeros((Bw, 6), dtype=torch.int32, device=dev)
        w["ftr"] = torch.zeros((Bw, 2860), dtype=torch.uint8, device=dev)
        w["score"] = torch.zeros((Bw, Tw), dtype=torch.int32, device=dev)
        # N > 1: two score buffers used in turn, so that the template scan of step i+1 never waits for the gather of step i
        w["score_alt"] = torch.zeros((Bw, Tw), dtype=torch.int32, device=dev) if world > 1 else None
        w["k"] = 0
        w["bidx"] = torch.zeros(Bw, dtype=torch.int32, device=dev)
        w["bdis"] = torch.zeros(Bw, dtype=torch.int32, device=dev)
        w["cmd"] = torch.zeros(Bw, dtype=torch.int32, device=dev)
        w["status"] = torch.zeros(Bw, dtype=torch.uint8, device=dev)
        w["gathered"] = torch.zeros((world * Bw, Tw), dtype=torch.int32, device=dev) if world > 1 else None
        w["gbest"] = torch.zeros(world * Bw, dtype=torch.int64, device=dev) if world > 1 else None
        w["outs"] = dict(seg_off=w["seg"].data_ptr(), ftr=w["ftr"].data_ptr(), score=w["score"].data_ptr(),
                         best_idx=w["bidx"].data_ptr(), best_dis=w["bdis"].data_ptr(), cmd=w["cmd"].data_ptr(),
                         status=w["status"].data_ptr())
        return w

    def step(w):
        h.set_bank_dev(w["bank"].data_ptr(), w["T"], 4096)
        if world > 1:
            # spch_recg on this rank's shard + the one exchange step of the path: NCCL all-gather of the u32
            # scores and the 8-byte argmin keys, on the communicator's own stream: it overlaps the next step entirely
            outs = dict(w["outs"])
            w["last_score"] = w["score_alt"] if (w["k"] & 1) else w["score"]
            outs["score"] = w["last_score"].data_ptr()
            w["k"] += 1
            h.recognise_dev_allgather(w["pcm"].data_ptr(), U, w["B"], N_LEN, gathered_score=w["gathered"].data_ptr(),
                                      gathered_best=w["gbest"].data_ptr(), **outs)
        else:
            h.recognise_dev(w["pcm"].data_ptr(), U, w["B"], N_LEN, **w["outs"])

    def timed_pass(w, steps, warmup, sampler=None):
        """W warm-up steps, then exactly K steps between barrier + synchronize, CUDA events on the launching stream,
        max over ranks; returns (ms per step, [(tag, ms)] kernel records of this rank, launches)"""
        for _ in range(max(warmup, 3)):
            step(w)
        if world > 1:
            h.comm_wait()
        barrier()
        l0 = h.launch_count()
        h.timing_enable(6 * steps + 8)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if sampler:
            sampler.mark()
        ev0.record(stream)
        for _ in range(steps):
            step(w)
        if world > 1:
            h.comm_wait()                               # the last step's gather is inside the timed region
        ev1.record(stream)
        barrier()
        if sampler:
            sampler.mark_end()
        ms_total = ev0.elapsed_time(ev1)
        launches = h.launch_count() - l0
        recs = h.timing_collect()
        h.timing_enable(0)
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, recs, launches

    def kernel_means(recs):
        per = {}
        for tag, ms in recs:
            per.setdefault(TAGS.get(tag, str(tag)), []).append(ms)
        return {k: float(np.mean(v)) for k, v in per.items()}

    w = make_workload(B, T)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, recs, launches = timed_pass(w, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if rank == 0 else None
    pcm, bank, seg, ftr, score = w["pcm"], w["bank"], w["seg"], w["ftr"], w.get("last_score", w["score"])
    bidx, bdis, cmd, status = w["bidx"], w["bdis"], w["cmd"], w["status"]

    # the exchange step's result must be the ranks' results in rank order
    gather_ok = None
    if world > 1:
        lo = rank * B
        key = (bdis.to(torch.int64) & 0xFFFFFFFF) << 32 | (bidx.to(torch.int64) & 0xFFFFFFFF)
        okrows = status == 0
        gather_ok = bool(torch.equal(w["gbest"][lo:lo + B], key) and
                         torch.equal(w["gathered"][lo:lo + B][okrows], score[okrows]))
        # and rank r's block on THIS rank equals what rank r computed: compare a checksum of every block
        chk = torch.stack([w["gbest"][r * B:(r + 1) * B].sum() for r in range(world)])
        mine = torch.zeros(world, dtype=torch.int64, device=dev)
        mine[rank] = key.sum()
        dist.all_reduce(mine)
        gather_ok = gather_ok and bool(torch.equal(chk, mine))

    # ---- per-utterance statistics of this rank's shard -------------------------------------------------
    frames_t = (ftr[:, 2].to(torch.int64) | (ftr[:, 3].to(torch.int64) << 8))
    seg64 = seg.to(torch.int64) & 0xFFFFFFFF
    ok = status == 0
    stats = torch.stack([frames_t.sum(), ok.sum(), ((seg64[:, 1] - seg64[:, 0]) * ok).sum()]).to(torch.float64)
    if world > 1:
        dist.all_reduce(stats)
    frames_total, ok_total, seg_samples = (float(x) for x in stats.tolist())

    # ---- e2e: host-buffer C-ABI call (pinned PCM in, cmd/dis/idx/status out) --------------------------
    # pinned memory on the GPU's own NUMA node (sr_host_alloc_dev), not torch's pin_memory(): the call is PCIe bound
    L = sr_b200.lib()
    pin_arr, pin_ptr = sr_b200.host_alloc_dev(local, B * U * 2)
    pin = torch.from_numpy(pin_arr.view(np.int16).reshape(B, U))
    pin.copy_(pcm)
    out_arr, out_ptr = sr_b200.host_alloc_dev(local, B * 16)
    o_idx = torch.from_numpy(out_arr[0:4 * B].view(np.int32))
    o_dis = torch.from_numpy(out_arr[4 * B:8 * B].view(np.int32))
    o_cmd = torch.from_numpy(out_arr[8 * B:12 * B].view(np.int32))
    o_st = torch.from_numpy(out_arr[12 * B:13 * B])
    numa = {"gpu_node": L.sr_device_numa_node(local), "thread_bound_to_node": bound_node,
            "pinned_pcm_node": L.sr_host_numa_node(C.c_void_p(pin_ptr))}
    he = sr_b200.Handle(local)
    he.set_bank_dev(bank.data_ptr(), T, 4096)
    ro = sr_b200.RecogOut(None, None, None, None, o_idx.data_ptr(), o_dis.data_ptr(), o_cmd.data_ptr(), o_st.data_ptr())

    def e2e_step():
        rc = L.sr_recognise_batch(he._h, C.c_void_p(pin_ptr), U, B, N_LEN, C.byref(ro))
        if rc != 0:
            raise RuntimeError(L.sr_last_error(he._h))

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()                                      # synchronous: returns after D2H + stream sync
    torch.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    e2e_equal = bool(torch.equal(o_cmd.to(dev), cmd) and torch.equal(o_dis.to(dev), bdis) and torch.equal(o_st.to(dev), status))
    tr_packed, tr_plain, tr_bytes = he.transport_stats()    # of the last step: chunks sent 12-bit packed / plain, bytes copied
    he.close()
    del pin, o_idx, o_dis, o_cmd, o_st
    sr_b200.host_free(pin_ptr)
    sr_b200.host_free(out_ptr)

    # ---- secondary passes the driver can see on the same line ---------------------------------------------
    # configs[3] (1 048 576 x 1 s, 50 templates, 131 072 utterances per GPU) with the same kernels, device resident
    config3 = None
    if world > 1 and args.config != 3 and not args.no_config3:
        B3, T3 = 131072, 50
        w3 = make_workload(B3, T3)
        k3 = max(3, min(args.steps, 10))
        ms3, recs3, _ = timed_pass(w3, k3, 3)
        ok3 = (w3["status"] == 0).sum().to(torch.float64)
        if world > 1:
            dist.all_reduce(ok3)
        config3 = {"workload": "configs[3]: %d x 1 s utterances over %d GPUs (%d per GPU), 12 MFCC, %d templates, "
                               "NCCL all-gather of u32 scores [%d,%d] + 8-byte argmin keys per rank" % (B3 * world, world, B3, T3, B3, T3),
                   "value": B3 * world / (ms3 * 1e-3), "unit": "utterances/s", "ms_per_step": ms3, "steps": k3,
                   "kernel_ms": kernel_means(recs3), "vad_ok_fraction": float(ok3.item()) / (B3 * world)}
        del w3
        torch.cuda.empty_cache()
    # configs[4]: 8 192 concurrent 5 s streams sharded over the GPUs, p50 / p99 per-utterance latency
    stream_line = None
    if not args.no_stream:
        stream_line = run_stream_shard(args, torch, dist, sr_b200, h, dev, stream, rank, world, local, bank, T)
        h.set_stream(stream.cuda_stream)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------
    kern_ms = kernel_means(recs)
    share = {k: v / max(sum(kern_ms.values()), 1e-9) for k, v in kern_ms.items()}
    frames_rank0 = float(frames_t.sum().item())
    seg_rank0 = float((((seg64[:, 1] - seg64[:, 0]) * ok).sum()).item())
    ok_rank0 = float(ok.sum().item())
    mfcc_bytes = 2.0 * (seg_rank0 + ok_rank0) + 24.0 * frames_rank0 + 4.0 * B     # 2*U_seg + 24*F + 4 per utterance
    peak, peak_src = peaks()
    ach = mfcc_bytes / (kern_ms.get("mfcc", float("nan")) * 1e-3) / 1e9
    # DRAM traffic / warp instructions per launch come from an ncu capture of THIS source tree (keyed on a hash of the
    # kernel sources: a capture of older kernels is reported as stale, never silently reused)
    traffic, traffic_src, tj = None, None, None
    sha = kernel_source_sha()
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        # "carried" lists later source hashes for which the capture still stands, each with the stated reason (a change
        # that moves no byte and no per-frame instruction); any other hash is stale
        carried = {c["sha"]: c["why"] for c in tj.get("carried", [])}
        if tj.get("kernel_source_sha") != sha and sha not in carried:
            traffic_src = "stale: profiles/r2_traffic.json was captured for kernel sources %s, this tree is %s" % (tj.get("kernel_source_sha"), sha)
            tj = None
        elif B == tj.get("batch") and T == tj.get("templates"):
            traffic = [v["dram_bytes_per_launch"] for k, v in tj["kernels"].items() if k.startswith("mfcc_kernel")][0]
            traffic_src = tj["source"]
            if sha in carried:
                traffic_src += " | captured for sources %s, carried to %s: %s" % (tj.get("kernel_source_sha"), sha, carried[sha])
    except Exception:
        tj = None
    vad_bytes = (2.0 * U + 24.0) * B                    # K0: 2*U read + 24 B written per utterance
    dtw_bytes = 24.0 * frames_rank0 + (4.0 + 4.0 * T) * B + T * 4096.0   # K2: features + scores (+ bank once)
    other = {}
    if kern_ms.get("vad"):
        a = vad_bytes / (kern_ms["vad"] * 1e-3) / 1e9
        other["vad_kernel"] = {"bound": "hbm", "achieved": a, "frac": a / peak, "algorithmic_bytes_per_launch": vad_bytes}
    if kern_ms.get("dtw"):
        a = dtw_bytes / (kern_ms["dtw"] * 1e-3) / 1e9
        other["dtw_kernel"] = {"bound": "hbm", "achieved": a, "frac": a / peak, "algorithmic_bytes_per_launch": dtw_bytes}
    roofline = {"bound": "hbm", "kernel": "mfcc_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src, "kernel_source_sha": sha,
                "peak_source": peak_src, "other_kernels": other,
                "algorithmic_bytes_per_launch": mfcc_bytes, "kernel_ms": kern_ms.get("mfcc"),
                "kernel_share_of_step": share.get("mfcc"),
                "note": "bit-exact fixed-point FFT: INT-issue bound by design (see int_issue and DESIGN.md), HBM is not what binds"}

    # ---- CPU baseline on a bounded sample + parity of the GPU results on that sample ----------------------
    cpu = None
    parity = None
    if not args.no_cpu:
        os.sched_setaffinity(0, orig_affinity)           # the CPU arm may use every core the job has, not just the GPU's socket
        cores = args.ref_procs if args.ref_procs > 0 else usable_cores()
        S = min(B, args.cpu_sample_per_core * cores)
        pcm_s = pcm[:S].cpu().numpy().view(np.uint16)
        bank_h = bank.cpu().numpy()
        secs, kind, out = cpu_reference_time(pcm_s, bank_h, T, cores, True)
        cpu = {"value": S / secs, "unit": "utterances/s", "cores": cores, "kind": kind,
               "sample": "first %d utterances of rank 0's batch, %d worker %s, wall = slowest worker"
                         % (S, cores, "processes" if kind == "reference" else "threads"),
               "mfcc_frames_per_s": float(out["frames"].sum() if "frames" in out else out["ftr"]["frm_num"].sum()) / secs}
        g = dict(seg_off=seg[:S].cpu().numpy().view(np.uint32).reshape(-1), best_idx=bidx[:S].cpu().numpy().view(np.uint32),
                 best_dis=bdis[:S].cpu().numpy().view(np.uint32), cmd=cmd[:S].cpu().numpy().view(np.uint32),
                 status=status[:S].cpu().numpy(), score=score[:S].cpu().numpy().view(np.uint32).reshape(-1))
        parity = all(np.array_equal(g[k], np.asarray(out[k]).reshape(-1)) for k in g)

    # integer-issue view of the dominant kernel: warp instructions per frame (ncu smsp__inst_executed.sum / frames,
    # a property of the SASS) x frames of this launch / live kernel time, against 4 schedulers x SMs x the SM clock
    # sampled under load. This, not HBM, is what bounds the bit-exact FFT.
    try:
        wipf = float(tj["mfcc_warp_inst_per_frame"])
        clk = float((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 0.0)
        n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        if clk > 0:
            ipeak = n_sms * 4 * clk * 1e6
            iach = wipf * frames_rank0 / (kern_ms["mfcc"] * 1e-3)
            roofline["int_issue"] = {"achieved": iach / 1e9, "peak": ipeak / 1e9, "unit": "G warp-inst/s", "frac": iach / ipeak,
                                     "warp_inst_per_frame": wipf,
                                     "note": "ALU and FMA-heavy (IMAD) pipes are half rate; pipe utilisation in " + str(tj.get("summary", "profiles/"))}
    except Exception:
        pass

    total_utts = B * world
    cfg_name = "configs[3]" if (B == 131072 and T == 50) else "configs[1]"
    line = {
        "metric": "utterances/s", "value": total_utts / (ms_step * 1e-3), "unit": "utterances/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32 fixed-point (s16 FFT, u32 energies)", "data": "synthetic",
        "config": {"workload": "%s: %d x %g s utterances (8 kHz u16), 12 MFCC, %d templates, per GPU; " % (cfg_name, B, U / 8000.0, T) +
                               "full spch_recg path: noise_atap+VAD -> get_mfcc(seg 0) -> dtw x T -> argmin",
                   "utterances_per_gpu": B, "samples_per_utterance": U, "templates": T, "geometry": "160/80/1024 (reference)",
                   "l2": "inputs (%.2f GB PCM per GPU) exceed the 126 MB L2; no flush needed" % (B * U * 2 / 1e9),
                   "multi_gpu": ("utterances sharded; one NCCL all-gather of u32 scores [B,T] + 8-byte argmin keys per step through the "
                                 "C-ABI (sr_recognise_batch_dev_allgather), on its own stream, overlapped with the next step") if world > 1 else "single GPU"},
        "mfcc_frames_per_s": frames_total / (ms_step * 1e-3),
        "dtw_pairs_per_s": ok_total * T / (ms_step * 1e-3),      # greedy walks per second (cells/s: bench.py --workload dtw)
        "vad_ok_fraction": ok_total / total_utts,
        "kernel_ms": kern_ms, "roofline": roofline, "cpu_baseline": cpu, "parity_vs_cpu_sample": parity,
        "e2e": {"value": total_utts / (e2e_ms * 1e-3), "unit": "utterances/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": tr_bytes if tr_bytes else B * U * 2, "d2h_bytes_per_step": B * 13, "matches_device_path": e2e_equal,
                "transport": {"chunks_packed_12bit": tr_packed, "chunks_plain_u16": tr_plain, "host_pcm_bytes_per_step": B * U * 2},
                "numa": numa, "call": "sr_recognise_batch (host pinned buffers from sr_host_alloc_dev)"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if gather_ok is not None:
        line["allgather_matches_rank_results"] = gather_ok
    if config3 is not None:
        line["config3"] = config3
    if stream_line is not None:
        line["config4_stream"] = stream_line
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def kernel_source_sha():
    """hash of the sources of the three kernels of the recognise step (K0 vad, K1 mfcc, K2 dtw) and of what they include:
    keys the ncu-derived numbers in profiles/r2_traffic.json (host-side files may change without invalidating a capture)"""
    import hashlib
    hsh = hashlib.sha256()
    d = os.path.join(ROOT, "stm32-speech-recognition_b200", "csrc")
    for name in ("sr_common.cuh", "sr_vad_core.cuh", "sr_vad.cu", "sr_mfcc.cu", "sr_dtw.cu", "sr_tables.cu", "sr_tables.h"):
        hsh.update(name.encode())
        hsh.update(open(os.path.join(d, name), "rb").read())
    return hsh.hexdigest()[:16]


def run_stream_shard(args, torch, dist, sr_b200, h, dev, stream, rank, world, local, bank, T):
    """BASELINE configs[4]: --streams concurrent 5 s streams (3 words each) sharded over the ranks' GPUs, fed in lock-step
    chunks of 100 ms and 10 ms from pinned host memory; latency = host time from handing over the chunk that completes the
    closing frame (sample end+879) to the results being visible on the host (sr_streams_push returning).
    Returns the dict for the JSON line on rank 0 (None elsewhere)."""
    S_all, Ls = args.streams, 40000
    s0, s1 = S_all * rank // world, S_all * (rank + 1) // world
    S = s1 - s0
    arr, ptr = sr_b200.host_alloc_dev(local, S * Ls * 2)
    pin = torch.from_numpy(arr.view(np.int16).reshape(S, Ls))
    gen = torch.empty((S, Ls), dtype=torch.int16, device=dev)
    sr_b200.synth_pcm_dev(gen.data_ptr(), S, Ls, 0x5EED5000 + s0, 3, stream.cuda_stream)
    pin.copy_(gen)
    del gen
    h.set_bank_dev(bank.data_ptr(), T, 4096)
    torch.cuda.synchronize(dev)
    h.use_own_stream()
    pool = sr_b200.StreamPool(h, S, Ls, N_LEN)
    out = {}
    for chunk in (800, 80):
        lat, n_events, t_all = [], 0, 0.0
        for rep in range(2):                                   # rep 0 = warm-up
            pool.reset()
            if world > 1:
                dist.barrier()
            lat, n_events, per_push = [], 0, []
            t_start = time.perf_counter()
            for n0 in range(0, Ls, chunk):
                t0 = time.perf_counter()
                ne = pool.push_raw(ptr + 2 * n0, chunk, Ls)   # returns when the events are in host memory
                dt = time.perf_counter() - t0
                if ne:
                    lat += [dt] * ne
                    n_events += ne
                    per_push.append((ne, dt * 1e3))
            t_all = time.perf_counter() - t_start
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, (lat, n_events, t_all))
            lat = [x for g in gathered for x in g[0]]
            n_events = sum(g[1] for g in gathered)
            t_all = max(g[2] for g in gathered)
        la = np.array(lat) * 1e3
        out["chunk_%d" % chunk] = {"chunk_ms": chunk / 8.0, "pushes": Ls // chunk, "events": n_events,
                                   "latency_ms_p50": float(np.percentile(la, 50)), "latency_ms_p99": float(np.percentile(la, 99)),
                                   "latency_ms_max": float(la.max()), "wall_s": t_all,
                                   "largest_pushes_events_ms": [[int(a), round(b, 3)] for a, b in sorted(per_push, reverse=True)[:4]],
                                   "realtime_factor": (S_all * Ls / 8000.0) / t_all, "utterances_per_s": n_events / t_all}
    seg, _ = pool.segments()
    pool.close()
    nchk = min(S, 128)
    hv = sr_b200.Handle(local)
    sub = np.ascontiguousarray(arr.view(np.uint16).reshape(S, Ls)[:nchk])
    ref = hv.vad(sub, hv.noise_atap(sub, N_LEN))
    hv.close()
    same = bool(np.array_equal(seg[:nchk], ref))
    del pin
    sr_b200.host_free(ptr)
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, same)
        same = all(flags)
    if rank != 0:
        return None
    return {"metric": "p50 per-utterance latency (streaming)", "value": out["chunk_800"]["latency_ms_p50"], "unit": "ms",
            "higher_is_better": False, "n_gpus": world,
            "workload": "configs[4]: %d concurrent 5 s streams (3 words each) sharded over %d GPU(s) (%d per GPU), %d templates, "
                        "lock-step chunks from pinned host memory" % (S_all, world, S_all // world, T),
            "results": out, "segments_equal_batch_vad_sample": same}


if __name__ == "__main__":
    main()
