#!/usr/bin/env python3
"""bench.py -- the driver's measurement contract for the svt-av1-psy B200 DSP tier.

  python bench.py --gpus N --steps K --warmup W                 B200 arm (libsvtav1_b200.so, T2 entry points)
  python bench.py --impl reference --gpus N --steps K --warmup W  reference arm: the reference's own kernels
                                                                (oracle/_ref, AVX2 intrinsics tier where it
                                                                builds without NASM, else C) on all host cores

A "step" = the hot-path DSP work of ONE 1920x1080 8-bit 4:2:0 frame at preset-8 / CRF-30 settings
(svt-av1-psy_b200/workload.py): open-loop ME (HME pyramid + 85-PU full-pel search, 2 references),
forward transform + quantize + inverse/reconstruction of every sample, CDEF search + apply, Wiener
statistics + filter.  Prints ONE JSON line (rank 0).

`value`  : frames/s with every input already resident in HBM (CUDA events on the launch stream).
`e2e`    : same metric through the host-buffer path: per step the source picture, residual and
           prediction are copied from pinned host memory and the ME results, quantised coefficients,
           CDEF costs, Wiener statistics and the filtered picture are read back, inside the timed region.
"""
import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json's metric, verbatim.  What is timed under that name is the tier's hot path only: one
# "frame" = ME + transform/quantise/inverse + CDEF + Wiener work of one 1080p preset-8 picture, NOT a full encode --
# `metric_scope` in the JSON line and config.workload say so, for both arms.
METRIC = "1080p preset-8 encoded frames/sec at 1/2/4/8 B200 vs reference AVX2 on host"
METRIC_SCOPE = "hot path only (SURVEY 8: ME + transform/quant/inverse + CDEF + Wiener of one 1080p preset-8 frame per step), not a full encode"
N_FRAME_SETS = 8  # rotated between steps: no step finds its inputs in L2, and (e2e) up to 8 frames are in flight
N_CALLS = 8       # len(FramePipeline.CALLS): the T2 entry points one frame goes through
# dram__bytes_read.sum + dram__bytes_write.sum of the call's dominant kernel, per launch, from the ncu --set full
# capture of this same command (profiles/README.md says which file); None = not captured for that call
NCU_DRAM_SOURCE = "profiles/r1_top_kernels_ncu_raw.csv (ncu --set full, one launch of the call's main kernel)"
NCU_DRAM_BYTES = {"cdef_search": 4995584 + 0,          # cdef_search_kernel: read + write (the mse output stays in L2)
                  "wiener_stats": 6377728 + 0,         # stats_mma_kernel
                  "txfm_trio": 999680 + 2682880 + 4361216 + 4520704 + 4305152 + 220000,  # the five class kernels (r + w)
                  "me_search": 7329024 + 12288 + 6755072}  # hme_fused_kernel (r + w) + fullpel_search_kernel  # hme_fused_kernel (r + w) + fullpel_search_kernel


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own kernels over the same work lists (oracle/ref_driver.c)
# ------------------------------------------------------------------------------------------------------
def aligned_zeros(n, dtype, align=64):
    """numpy array whose data pointer is `align`-byte aligned (the AVX2 kernels use aligned stores)"""
    isz = np.dtype(dtype).itemsize
    raw = np.zeros(n * isz + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * isz].view(dtype)


class RefFrame:
    def __init__(self, wl, ref):
        from oracle import support as me_np
        self.wl, self.ref = wl, ref
        W, H = wl.width, wl.height
        self.cur_pyr = me_np.build_pyramid_np(wl.cur[0], W, H, wl.me_shapes)
        self.ref_pyrs = [me_np.build_pyramid_np(r[0], W, H, wl.me_shapes) for r in wl.refs]
        self.cur_desc = me_np.ref_pic_desc(self.cur_pyr, wl.me_shapes)
        self.ref_descs = (me_np.RefMePicture * wl.n_refs)(*[me_np.ref_pic_desc(p, wl.me_shapes) for p in self.ref_pyrs])
        self.prm = (me_np.RefMeParams * wl.n_refs)()
        for i, p in enumerate(wl.me_params):
            for k, v in p.items():
                setattr(self.prm[i], k, v)
        nb = ((W + 63) // 64) * ((H + 63) // 64)
        self.me_sad = np.zeros((wl.n_refs, nb, 85), np.uint32)
        self.me_mv = np.zeros_like(self.me_sad)
        self.me_c = np.zeros((wl.n_refs, nb, 2), np.int16)
        self.me_hs = np.zeros((wl.n_refs, nb), np.uint64)
        self.cur_flat = np.concatenate([p.reshape(-1) for p in wl.cur])
        res = np.concatenate([p.reshape(-1) for p in wl.residual])
        self.residual = aligned_zeros(res.size, np.int16)
        self.residual[:] = res
        _, n_pad = wl.padded_offsets()
        self.pred = self._pad_planes(wl.pred)
        self.recon = np.zeros(n_pad, np.uint8)
        self.cdef_out = np.zeros(n_pad, np.uint8)
        self.final = np.zeros(n_pad, np.uint8)
        self.coeff = aligned_zeros(wl.n_coeffs, np.int32)
        self.q = aligned_zeros(wl.n_coeffs, np.int32)
        self.dq = aligned_zeros(wl.n_coeffs, np.int32)
        self.eobs = np.zeros(len(wl.quant_items), np.uint16)
        self.fwd = np.ascontiguousarray(wl.fwd_items)
        self.inv = np.ascontiguousarray(wl.inv_items)
        self.qi = np.ascontiguousarray(wl.quant_items)
        self.mse = np.zeros((2, nb, len(wl.cdef_str_y)), np.uint64)
        self.dirs = np.zeros((nb, 64), np.uint8)
        self.vars = np.zeros((nb, 64), np.int32)
        self.M = np.zeros((len(wl.stats_items), 49), np.int64)
        self.Hm = np.zeros((len(wl.stats_items), 2401), np.int64)
        for f in ("ref_me_picture", "ref_fwd_txfm_batch", "ref_quant_batch", "ref_inv_txfm_batch_8bit", "ref_cdef_search_frame",
                  "ref_cdef_apply_frame", "ref_compute_stats_batch", "ref_wiener_units_8bit"):
            getattr(ref, f).restype = None

    def _pad_planes(self, planes):
        wl = self.wl
        off, n = wl.padded_offsets()
        buf = np.zeros(n, np.uint8)
        for p in range(3):
            th, st = wl.padded_shape(p)
            w, h = wl.plane_dims[p]
            buf[off[p]:off[p] + th * st].reshape(th, st)[:, :w + 2 * wl.PAD] = np.pad(planes[p], wl.PAD, mode="edge")
        return buf

    def _extend(self, buf):
        wl = self.wl
        off, _ = wl.padded_offsets()
        for p in range(3):
            th, st = wl.padded_shape(p)
            w, h = wl.plane_dims[p]
            v = buf[off[p]:off[p] + th * st].reshape(th, st)
            v[:, :w + 2 * wl.PAD] = np.pad(v[wl.PAD:wl.PAD + h, wl.PAD:wl.PAD + w], wl.PAD, mode="edge")

    def _cdef_frame(self):
        from oracle import support as me_np
        wl = self.wl
        off, _ = wl.padded_offsets()
        soff, _ = wl.flat_offsets()
        f = me_np.RefCdefFrame()
        ptrs = []
        for p in range(3):
            th, st = wl.padded_shape(p)
            ptrs.append(self.recon.ctypes.data + off[p] + wl.PAD * st + wl.PAD)
        f.recon_y, f.recon_cb, f.recon_cr = ptrs
        f.src_y, f.src_cb, f.src_cr = [self.cur_flat.ctypes.data + soff[p] for p in range(3)]
        f.recon_stride_y, f.recon_stride_c = wl.padded_shape(0)[1], wl.padded_shape(1)[1]
        f.src_stride_y, f.src_stride_c = wl.plane_dims[0][0], wl.plane_dims[1][0]
        f.width, f.height, f.bit_depth, f.damping, f.subsampling_factor = wl.width, wl.height, 8, wl.cdef_damping, wl.cdef_subsampling
        return f

    def step(self):
        wl, ref = self.wl, self.ref
        P = lambda a, o=0: ct.c_void_p(a.ctypes.data + o)  # noqa: E731
        # ME (the numpy pyramid build is input preparation and stays outside, like the resident refs)
        ref.ref_me_picture(ct.byref(self.cur_desc), self.ref_descs, self.prm, wl.n_refs, P(self.me_sad), P(self.me_mv), P(self.me_c), P(self.me_hs))
        ref.ref_fwd_txfm_batch(P(self.residual), P(self.coeff), P(self.fwd), len(self.fwd))
        ref.ref_quant_batch(P(self.coeff), P(self.q), P(self.dq), P(wl.scan_table), P(wl.iscan_table), P(wl.qm_table), P(self.qi), len(self.qi), P(self.eobs))
        ref.ref_inv_txfm_batch_8bit(P(self.dq), P(self.pred), P(self.recon), P(self.inv), len(self.inv))
        f = self._cdef_frame()
        ref.ref_cdef_search_frame(ct.byref(f), P(wl.skip8x8), P(wl.cdef_str_y), P(wl.cdef_str_uv), len(wl.cdef_str_y), P(self.mse), P(self.dirs),
                                  P(self.vars))
        np.copyto(self.cdef_out, self.recon)
        off, _ = wl.padded_offsets()
        outs = [self.cdef_out.ctypes.data + off[p] + wl.PAD * wl.padded_shape(p)[1] + wl.PAD for p in range(3)]
        ref.ref_cdef_apply_frame(ct.byref(f), P(wl.skip8x8), P(wl.cdef_fb_idx), P(wl.cdef_apply_y), P(wl.cdef_apply_uv), ct.c_void_p(outs[0]),
                                 ct.c_void_p(outs[1]), ct.c_void_p(outs[2]), wl.padded_shape(0)[1], wl.padded_shape(1)[1])
        self._extend(self.cdef_out)
        ref.ref_compute_stats_batch(P(self.cdef_out), P(self.cur_flat), P(np.ascontiguousarray(wl.stats_items)), len(wl.stats_items), P(self.M),
                                    P(self.Hm))
        ref.ref_wiener_units_8bit(P(self.cdef_out), P(self.final), P(np.ascontiguousarray(wl.wiener_units)), len(wl.wiener_units))


def load_reference():
    import oracle
    if oracle.ref is None:
        return None, "port", 0
    tier = oracle.ref.ref_set_tier(1)
    return oracle.ref, ("reference" if tier == 1 else "reference"), tier


def time_reference(wl, steps, warmup):
    ref, kind, tier = load_reference()
    if ref is None:
        return None
    fr = RefFrame(wl, ref)
    cores = ref.ref_num_threads()
    for _ in range(warmup):
        fr.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        fr.step()
    dt = time.perf_counter() - t0
    return dict(fps=steps / dt, ms=1e3 * dt / steps, cores=cores, kind=kind,
                tier="avx2-intrinsics (inverse transform: C, no NASM)" if tier == 1 else "c", frame=fr)


# ------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi) during the timed region
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML in-process (a sample every ~0.5 ms --
    the timed region is only tens of milliseconds long), nvidia-smi as the fallback."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]  # nvmlClocksEventReason{HwSlowdown,HwThermalSlowdown,SwThermalSlowdown,SwPowerCap}

    def __init__(self, index, bus_id=None):
        self.samples, self.index, self.stop_flag, self.th = [], index, False, None
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id) if bus_id else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _run(self):
        while not self.stop_flag:
            if self.nvml is not None:
                try:
                    mhz = self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
                    try:
                        mask = self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                    except Exception:
                        mask = self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                    self.samples.append([str(mhz), str(self.max_mhz)] + ["Active" if mask & b else "Not Active" for b in self.BITS])
                except Exception:
                    pass
                time.sleep(0.0005)
                continue
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        reasons = sorted({self.NAMES[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--check", action="store_true", help="compare one frame of B200 output with the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--one-stream", action="store_true", help="all frames on one compute stream (no frame-level overlap)")
    ap.add_argument("--streams", type=int, default=4, help="compute streams that consecutive frames alternate between (1, 2 or 4)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying one CUDA graph per step")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3)

    import svt_av1_psy_b200  # noqa: F401  (ImportError = library not built: there is no fallback)
    from svt_av1_psy_b200.workload import FrameWorkload

    config = {"workload": "configs[1]: 1920x1080 8-bit 4:2:0 synthetic, preset 8 CRF 30 hot path (ME 2 refs + TX + CDEF + Wiener), 1 frame/step",
              "width": args.width, "height": args.height, "frame_sets": N_FRAME_SETS,
              "l2": "steps rotate over %d distinct frame sets (~65 MB each, >126 MB L2 in total)" % N_FRAME_SETS,
              "parallelism": "frame-parallel x%d (no data-path collective; recon exchange = all_gather of the filtered frame)" % world,
              "overlap": "ME (source pictures only) on a side stream, concurrent with the transform->CDEF->restoration chain of the same step",
              "streams": "1 compute stream" if args.one_stream else "%d compute streams: consecutive (independent) frames alternate between them" % args.streams,
              "launch": "eager" if args.no_graph else "one CUDA graph replay per step (the step's kernel launches captured once per frame set)"}

    if args.impl == "reference":
        if rank != 0:
            return
        wl = FrameWorkload(args.width, args.height)
        steps = min(args.steps, 5)
        t = time_reference(wl, steps, min(warmup, 1))
        if t is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsvtav1_ref.so is not built"}))
            return
        out = {"impl": "reference", "metric": METRIC, "metric_scope": METRIC_SCOPE, "value": round(t["fps"], 3), "unit": "frames/s",
               "n_gpus": args.gpus, "steps": steps,
               "warmup": min(warmup, 1), "ms_per_step": round(t["ms"], 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "u8", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": round(t["fps"], 3), "unit": "frames/s", "cores": t["cores"], "kind": "reference",
                                "sample": "%d full frames, tier %s" % (steps, t["tier"])},
               "e2e": {"value": round(t["fps"], 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    from svt_av1_psy_b200 import dsp
    from svt_av1_psy_b200.pipeline import FramePipeline
    assert len(FramePipeline.CALLS) == N_CALLS
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dsp.init(local_rank)
    wl0 = FrameWorkload(args.width, args.height, seed=20260923 + 17 * rank * N_FRAME_SETS)
    sets = [FramePipeline(wl0 if i == 0 else wl0.with_seed(20260923 + 17 * (rank * N_FRAME_SETS + i)), torch) for i in range(N_FRAME_SETS)]
    stream = torch.cuda.Stream()
    # consecutive frames alternate between two compute streams: independent pictures in flight at once, as in
    # the encoder's picture-parallel pipeline; frame set k always runs on stream k % 2
    n_streams = 1 if args.one_stream else args.streams
    assert n_streams in (1, 2, 4, 8), "--streams must divide the %d frame sets" % N_FRAME_SETS
    streams = [stream] + [torch.cuda.Stream() for _ in range(n_streams - 1)]
    gathered = [torch.empty((world,) + tuple(sets[0].final.shape), dtype=torch.uint8, device="cuda") for _ in range(N_FRAME_SETS)] if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphs = [None] * N_FRAME_SETS

    def enqueue_step(i, events=None):
        """one frame of hot-path work on the current stream: replay of the frame set's CUDA graph (the ~35
        kernel launches of a step captured once; same kernels, same work) or, for per-call timing, eager"""
        g = graphs[i % N_FRAME_SETS]
        if g is not None and events is None:
            g.replay()
        else:
            sets[i % N_FRAME_SETS].step(events)

    def capture_graphs():
        for k, fp in enumerate(sets):
            # one eager step on the stream the graph will be captured on: every lazily allocated library
            # workspace (they are per stream) exists before the capture, which must not allocate
            with torch.cuda.stream(streams[k % n_streams]):
                fp.step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[k % n_streams]):
                fp.step()
            graphs[k] = g

    def run(n, e2e, stage_acc=None):
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(N_CALLS + 1)] for _ in range(n)] if stage_acc is not None else None
        start, end, tail = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
        start.record(stream)
        for x in streams[1:]:
            x.wait_event(start)
        for i in range(n):
            st = stream if ev else streams[i % n_streams]  # the per-call profile runs strictly serially
            with torch.cuda.stream(st):
                fp = sets[i % N_FRAME_SETS]
                if e2e:
                    fp.load_inputs()
                enqueue_step(i, ev[i] if ev else None)
                if world > 1:  # reconstructed-reference exchange (the path's one real collective)
                    dist.all_gather_into_tensor(gathered[i % N_FRAME_SETS].view(-1), fp.final)
                if e2e:
                    fp.read_outputs()
        for x in streams[1:]:
            tail.record(x)
            stream.wait_event(tail)
        end.record(stream)
        torch.cuda.synchronize()
        if stage_acc is not None:
            for i in range(n):
                for k in range(N_CALLS):
                    stage_acc[k] += ev[i][k].elapsed_time(ev[i][k + 1])
        return start.elapsed_time(end)

    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def run_e2e(n):
        """host buffers -> device -> host, every step; the three phases of consecutive frames overlap on
        three streams (copy-in / compute / copy-out), ordered with events; a frame set is re-used only
        after its previous results have been read back."""
        ev_in = [torch.cuda.Event() for _ in range(n)]
        ev_done = [torch.cuda.Event() for _ in range(n)]
        ev_out = [torch.cuda.Event() for _ in range(n)]
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        s_in.wait_event(start)
        for i in range(n):
            fp = sets[i % N_FRAME_SETS]
            with torch.cuda.stream(s_in):
                if i >= N_FRAME_SETS:
                    s_in.wait_event(ev_out[i - N_FRAME_SETS])
                fp.load_inputs()
                ev_in[i].record(s_in)
            cs = streams[i % n_streams]
            with torch.cuda.stream(cs):
                cs.wait_event(ev_in[i])
                enqueue_step(i)
                if world > 1:
                    dist.all_gather_into_tensor(gathered[i % N_FRAME_SETS].view(-1), fp.final)
                ev_done[i].record(cs)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[i])
                fp.read_outputs()
                ev_out[i].record(s_out)
        stream.wait_event(ev_out[n - 1])
        end.record(stream)
        torch.cuda.synchronize()
        return start.elapsed_time(end)

    if args.check and rank == 0:
        check_against_reference(sets[0], torch)

    # ---- resident-input timing ---------------------------------------------------------------------------
    run(warmup, False)
    barrier()
    # per-call profile: eager launches with an event around every T2 call (not part of the timed value)
    l0 = dsp.launch_count()
    call_ms = [0.0] * N_CALLS
    ms_eager = run(args.steps, False, call_ms)
    launches = dsp.launch_count() - l0
    if not args.no_graph:
        capture_graphs()
        run(warmup, False)
    barrier()
    bus_id = None
    try:  # NVML enumerates physical devices: address this rank's GPU by PCI id, not by (visible) index
        pr = torch.cuda.get_device_properties(local_rank)
        bus_id = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        bus_id = None
    sampler = ClockSampler(local_rank, bus_id)
    sampler.start()
    ms = run(args.steps, False)
    barrier()
    # ---- end-to-end timing -----------------------------------------------------------------------------------
    run_e2e(warmup)
    barrier()
    ms_e2e = run_e2e(args.steps)
    clocks = sampler.stop()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fps = world * args.steps / (ms / 1e3)
    fps_e2e = world * args.steps / (ms_e2e / 1e3)
    alg = wl0.algorithmic_bytes()
    call_ms = [x / args.steps for x in call_ms]
    calls = FramePipeline.CALLS
    names = list(FramePipeline.STAGES)
    stage_ms = [sum(call_ms[i] for i, c in enumerate(calls) if c[1] == st) for st in names]
    dom = int(np.argmax(call_ms))
    dom_name, _, dom_kernels = calls[dom]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    src = "MEASURED_PEAKS.json" if peaks else "fallback of /opt/skills/guides/B200_PROFILING.md"
    if dom_name == "wiener_stats":  # the one dense contraction of the path: exact f16 MMA on the tensor cores
        flops = 2.0 * wl0.wiener_stats_macs()
        peak = float(peaks.get("bf16_tflops", peaks.get("dense_bf16_tflops", 2250.0)))
        ach = flops / (call_ms[dom] / 1e3) / 1e12
        roofline = {"bound": "tensor", "achieved": round(ach, 3), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 5),
                    "algorithmic_flops_per_call": flops}
    else:
        peak = float(peaks.get("hbm_gbs", 6650.0))
        ach = alg[dom_name] / (call_ms[dom] / 1e3) / 1e9
        roofline = {"bound": "hbm", "achieved": round(ach, 2), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 5),
                    "algorithmic_bytes_per_call": alg[dom_name],
                    "note": "integer kernel working out of shared-memory tiles: instruction/latency bound, not HBM bound (SURVEY 8d)"}
    roofline.update({"call": dom_name, "kernel": dom_kernels, "ms_per_call": round(call_ms[dom], 4), "peak_source": src,
                     "traffic": NCU_DRAM_BYTES.get(dom_name), "traffic_source": NCU_DRAM_SOURCE if dom_name in NCU_DRAM_BYTES else None})
    out = {"metric": METRIC, "value": round(fps, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
           "ms_per_step": round(ms / args.steps, 4), "metric_scope": METRIC_SCOPE, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
           "data": "synthetic", "config": config, "clocks": clocks,
           "e2e": {"value": round(fps_e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": int(sets[0].h2d_bytes), "d2h_bytes_per_step": int(sets[0].d2h_bytes),
                   "ms_per_step": round(ms_e2e / args.steps, 4)},
           "gpu_launches": int(launches), "eager_ms_per_step": round(ms_eager / args.steps, 4), "roofline": roofline,
           "stages_ms": {n: round(v, 4) for n, v in zip(names, stage_ms)},
           "calls_ms": {c[0]: round(v, 4) for c, v in zip(calls, call_ms)},
           "calls_algorithmic_gbs": {c[0]: round(alg[c[0]] / (v / 1e3) / 1e9, 2) for c, v in zip(calls, call_ms) if v > 0}}
    if world == 1 and not args.no_cpu_baseline:
        t = time_reference(wl0, 2, 1)
        if t is not None:
            out["cpu_baseline"] = {"value": round(t["fps"], 3), "unit": "frames/s", "cores": t["cores"], "kind": "reference",
                                   "sample": "2 full frames of the same workload, tier %s" % t["tier"]}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def check_against_reference(fp, torch):
    """one frame through both arms, every output compared bit for bit"""
    ref, _, _ = load_reference()
    assert ref is not None, "oracle/_ref missing"
    ref.ref_set_tier(0)
    fr = RefFrame(fp.wl, ref)
    fr.step()
    fp.load_inputs()
    fp.step()
    torch.cuda.synchronize()
    cmp = [("me_sad", fp.me_sad, fr.me_sad), ("me_mv", fp.me_mv, fr.me_mv), ("hme_centre", fp.me_centre, fr.me_c),
           ("qcoeff", fp.qcoeff, fr.q), ("dqcoeff", fp.dqcoeff, fr.dq), ("eob", fp.eobs, fr.eobs), ("recon", fp.recon, fr.recon),
           ("cdef_mse", fp.cdef_mse, fr.mse), ("cdef_dir", fp.cdef_dir, fr.dirs), ("cdef_out", fp.cdef_out, fr.cdef_out), ("wiener_M", fp.M, fr.M),
           ("wiener_H", fp.Hm, fr.Hm), ("final", fp.final, fr.final)]
    bad = []

    def compare(items):
        for name, a, b in items:
            a = a.cpu().numpy()
            if not np.array_equal(a.view(np.uint8).reshape(-1), np.ascontiguousarray(b).view(np.uint8).reshape(-1)):
                bad.append(name)

    compare(cmp)
    # the frame step uses the fused transform call; the same chain as three separate calls (which also
    # materialises the forward coefficients) must give the same answers
    s = torch.cuda.current_stream().cuda_stream
    for t in (fp.qcoeff, fp.dqcoeff, fp.eobs, fp.recon):
        t.zero_()
    fp.call_fwd_txfm(s)
    fp.call_quant(s)
    fp.call_inv_txfm(s)
    torch.cuda.synchronize()
    split = [("coeff", fp.coeff, fr.coeff), ("qcoeff/3-call", fp.qcoeff, fr.q), ("dqcoeff/3-call", fp.dqcoeff, fr.dq),
             ("eob/3-call", fp.eobs, fr.eobs), ("recon/3-call", fp.recon, fr.recon)]
    compare(split)
    cmp = cmp + split
    if bad:
        raise SystemExit("PARITY FAILURE vs reference: " + ", ".join(bad))
    print("parity vs reference C tier: all %d outputs bit-exact" % len(cmp), file=sys.stderr)


if __name__ == "__main__":
    main()
