#!/usr/bin/env python
"""Benchmarks of the MLIC++ forward path on B200 (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W                  # headline: MLICPP_L forward @1920x1088 (configs[1])
    python bench.py --config sd_decode | vbr_sweep | 4k_bands ...  # configs[2], [3] and the row-band half of [4]
    python bench.py --size 4k ...                                  # configs[4], per-image sharding
    python bench.py --impl reference ...                           # the reference's CPU forward (oracle port) on the host cores

One process per GPU (torchrun for N > 1).  The default workloads shard by image (weak scaling, no data-path collective; one
all-reduce carries the aggregate rate / distortion sums); `4k_bands` splits ONE 3840x2176 image into row bands (strong scaling:
NCCL P2P halos + an all-gather of the latent).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MP_1080 = 1920 * 1088 / 1e6
# name -> model, H, W, nominal MP per image, algorithmic FLOP per image, default images per GPU, metric, workload text
CONFIGS = {
    "forward": dict(model="MLICPP_L", H=1088, W=1920, mp=MP_1080, flop=1.753e12, batch=32,
                    metric="megapixels/sec MLICPP_L forward @1920x1088",
                    workload="MLICPP_L forward 1920x1088 (BASELINE configs[1])"),
    "forward_4k": dict(model="MLICPP_L", H=2176, W=3840, mp=3840 * 2160 / 1e6, flop=4 * 1.753e12, batch=8,
                       metric="megapixels/sec MLICPP_L forward @3840x2160",
                       workload="MLICPP_L forward 3840x2160 padded to 3840x2176, sharded by image (BASELINE configs[4])"),
    "sd_decode": dict(model="MLICPP_M_SMALL_DEC", H=1088, W=1920, mp=MP_1080, flop=0.371e12, batch=32,
                      metric="megapixels/sec MLICPP_M_SMALL_DEC decode-side walk @1920x1088",
                      workload="MLICPP_M_SMALL_DEC net_decoder_forward (h_s + entropy model + depthwise-separable g_s) 1920x1088 (BASELINE configs[2])"),
    "vbr_sweep": dict(model="MLICPP_L_VBR", H=1088, W=1920, mp=MP_1080, flop=None, batch=8,
                      metric="megapixels/sec MLICPP_L_VBR compress-path symbol / CDF-index generation @1920x1088, 6 rates",
                      workload="MLICPP_L_VBR compress-path network walk (symbols + CDF indexes) at each of the 6 rate levels, 1920x1088 "
                               "(BASELINE configs[3]); every image counts once per rate"),
    "4k_bands": dict(model="MLICPP_L", H=2176, W=3840, mp=3840 * 2160 / 1e6, flop=4 * 1.753e12, batch=1,
                     metric="megapixels/sec MLICPP_L forward of ONE 3840x2160 image split into row bands",
                     workload="MLICPP_L forward, one 3840x2160 image (padded to 3840x2176) in row bands over the ranks: image-row halos by NCCL "
                              "P2P, all-gather of y, replicated entropy model, g_s on band + 7 latent halo rows (BASELINE configs[4], row-band half)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sust=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def seeded_model(name, device=None):
    """This repo's model with the seeded benchmark weights (oracle.weights only generates the numbers)."""
    import mlic_b200
    from oracle import weights
    net = mlic_b200.get_model(name)
    net.load_state_dict(weights.seeded_state_dict(net.state_dict(), 1234))
    net.update(force=True)
    return net.to(device) if device is not None else net


# ---------------------------------------------------------------------------------------------------- CPU legs (oracle only)
def cpu_oracle(model):
    """The CPU restatement of the reference with the same seeded weights, built from the committed state_dict template:
    nothing of the product package (mlic_b200, its .so) is imported on this path."""
    from oracle import mlic_oracle, weights
    return mlic_oracle.Oracle(model, weights.reference_state_dict(model, 1234))


def cpu_step(orc, cfg_name, x):
    """One unit of the workload on the CPU -> megapixels processed."""
    h, w = x.shape[-2:]
    if cfg_name == "sd_decode":
        orc.decoder_forward(x)
        return h * w / 1e6
    if cfg_name == "vbr_sweep":
        for s in range(6):
            orc.compress_symbols(x, s=s)
        return 6 * h * w / 1e6
    orc.forward(x)
    return h * w / 1e6


def cpu_baseline_mps(cfg_name, cfg, threads, min_seconds=15.0, max_runs=6):
    from oracle import weights
    torch.set_num_threads(threads)
    orc = cpu_oracle(cfg["model"])
    H, W = (1088, 1920) if cfg_name in ("forward_4k", "4k_bands") else (cfg["H"], cfg["W"])     # bounded sample: one 1080p image
    x = weights.synthetic_image(1, H, W, seed=2024, kind="rand")
    times, mp = [], 0.0
    with torch.no_grad():
        while len(times) < 2 or (sum(times) < min_seconds and len(times) < max_runs):
            t0 = time.perf_counter()
            mp = cpu_step(orc, cfg_name, x)
            times.append(time.perf_counter() - t0)
    return mp / min(times), times, (H, W)


def run_reference(args, cfg_name, cfg, rank):
    """--impl reference: the reference algorithm on the host cores (CPU port in oracle/, all threads), same config / metric / unit.
    Each step is one whole image of the workload (1920x1088; for the 4K configs a 1920x1088 image is the bounded sample)."""
    if rank != 0:
        return
    from oracle import weights
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    orc = cpu_oracle(cfg["model"])
    H, W = (1088, 1920) if cfg_name in ("forward_4k", "4k_bands") else (cfg["H"], cfg["W"])
    if os.environ.get("MLIC_BENCH_REF_SMALL"):           # the CPU test suite checks the contract line, not the host's speed
        H, W = 128, 192
    x = weights.synthetic_image(1, H, W, seed=2024, kind="rand")
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    # keep the whole run within a few minutes on slow hosts: probe one step, then cap the step count
    with torch.no_grad():
        t0 = time.perf_counter()
        mp = cpu_step(orc, cfg_name, x)
        probe = time.perf_counter() - t0
        budget = 240.0
        if (steps + warmup) * probe > budget:
            warmup = 0
            steps = max(1, int(budget / probe))
        for _ in range(max(warmup - 1, 0)):
            cpu_step(orc, cfg_name, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            mp = cpu_step(orc, cfg_name, x)
        dt = time.perf_counter() - t0
    mps = steps * mp / dt
    sample = (f"1 image {W}x{H} per step" + (" (a 1920x1088 image stands for the 4K workload)" if (H, W) != (cfg["H"], cfg["W"]) else "") +
              f", fp32, torch CPU (oracle port of the reference), {threads} threads, {steps} timed steps")
    line = {"impl": "reference", "metric": cfg["metric"], "value": mps, "unit": "MP/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
            "scaling": "strong" if cfg_name == "4k_bands" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "sample": sample, "weights": "random-init seed 1234 (oracle/shapes template)"},
            "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": mps, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="forward", choices=["forward", "sd_decode", "vbr_sweep", "4k_bands"],
                    help="forward: MLICPP_L forward (configs[1], the headline); sd_decode: MLICPP_M_SMALL_DEC decode-side walk (configs[2]); "
                         "vbr_sweep: MLICPP_L_VBR symbol generation at 6 rates (configs[3]); 4k_bands: one 4K image in row bands over the ranks")
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: per config)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--size", default="1080p", choices=["1080p", "4k"],
                    help="forward only. 4k: 3840x2160 padded to 3840x2176 as the reference pads (configs[4], per-image sharding)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    cfg_name = "forward_4k" if (args.config == "forward" and args.size == "4k") else args.config
    cfg = CONFIGS[cfg_name]
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg_name, cfg, rank)
        return
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: NCCL (version banner) and other native libraries write to file descriptor 1 directly,
    # so everything but that line is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    H, W, MP_IMG = cfg["H"], cfg["W"], cfg["mp"]
    B = args.batch or cfg["batch"]

    import torch.distributed as dist
    if world > 1 and os.environ.get("MLIC_BIND_CPUS", "1") != "0":
        from mlic_b200.dist import bind_to_gpu_cpus
        bind_to_gpu_cpus(local)                # host staging memory next to this rank's GPU (matters for `e2e` on two-socket boxes)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"          # NCCL's version banner goes to stdout: keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    from mlic_b200 import _lib
    from oracle import weights                  # seeded weights / synthetic images only (outside every timed region)
    net = seeded_model(cfg["model"], dev).set_precision(args.precision)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the step of each workload: dev_step() on device-resident inputs, host_step() through host buffers (e2e)
    bands = cfg_name == "4k_bands"
    if bands:
        from mlic_b200.dist import EngineStages, forward_row_bands, shard_range
        B = 1
        x_full = weights.synthetic_image(1, H, W, seed=2024, kind="rand")
        lo, hi = shard_range(H // 16, rank, world)
        x_host = x_full[:, :, 16 * lo:16 * hi].contiguous().pin_memory()
        x = x_host.to(dev, non_blocking=True)
        st = EngineStages(net)
        units_per_step = MP_IMG                  # the ONE image, whatever the rank count (strong scaling)

        def dev_step():
            return forward_row_bands(st, x, H // 16, rank, world)

        pin = {}

        def to_pinned(key, t):
            if key not in pin:
                pin[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            pin[key].copy_(t, non_blocking=True)
            return pin[key]

        def host_step():           # band of x up, this rank's x_hat rows (and, on rank 0, the likelihoods) down
            o = forward_row_bands(st, x_host.to(dev, non_blocking=True), H // 16, rank, world)
            res = [to_pinned("x_hat", o["x_hat_band"])]
            if rank == 0:
                res += [to_pinned("yl", o["likelihoods"]["y_likelihoods"]), to_pinned("zl", o["likelihoods"]["z_likelihoods"])]
            torch.cuda.current_stream().synchronize()
            return res
        h2d = x_host.numel() * 4
    else:
        x_host = weights.synthetic_image(B, H, W, seed=2024 + rank * B, kind="rand").pin_memory()
        x = x_host.to(dev, non_blocking=True)
        units_per_step = world * B * MP_IMG * (6 if cfg_name == "vbr_sweep" else 1)
        if cfg_name == "sd_decode":
            def dev_step():
                return net.net_decoder_forward(x)

            pin = {}

            def host_step():       # the decode-side walk reads only the SHAPE of x (mlicpp.py:380-394): nothing goes up, x_hat comes home
                xh = net.net_decoder_forward(x)
                if "x_hat" not in pin:
                    pin["x_hat"] = torch.empty(xh.shape, dtype=xh.dtype, pin_memory=True)
                pin["x_hat"].copy_(xh, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return pin["x_hat"]
            h2d = 0
        elif cfg_name == "vbr_sweep":
            def dev_step():
                return [net._run(_lib.MODE_COMPRESS, x, B, H, W, net._scale(s, 0, True)) for s in range(6)]

            def host_step():         # host x in, pinned host symbols / indexes / z symbols (+ x_hat) out, per rate
                return [net._run(_lib.MODE_COMPRESS, x_host, B, H, W, net._scale(s, 0, True)) for s in range(6)]
            h2d = 6 * x_host.numel() * 4
        else:
            def dev_step():
                return net(x)

            def host_step():
                return net(x_host)
            h2d = x_host.numel() * 4

    for _ in range(args.warmup):
        out = dev_step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = dev_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    launches = net.last_launch_count * args.steps * (6 if cfg_name == "vbr_sweep" else 1)
    # roofline pass: the same K steps again with a CUDA-event pair around every tcgen05 launch (on the launch stream).  Kept out
    # of the timed region above because ~400 event pairs per step cost the step itself ~4 % (measured: 37.8 vs 36.1 ms at 8 images).
    net.set_profile(True)
    net.profile_read(reset=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    p0.record()
    for _ in range(args.steps):
        out = dev_step()
    p1.record()
    barrier()
    ms_prof = p0.elapsed_time(p1)
    tc_ms, tc_flops, tc_launches = net.profile_read(reset=False)
    top_ms, top_flops, top_n = net.profile_read_top(reset=True)
    net.profile_read(reset=True)
    net.set_profile(False)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = units_per_step * args.steps / (ms / 1e3)

    # aggregate rate / distortion of the forward workloads (the only collective on the batch-sharded path: 3 numbers)
    quality = None
    if cfg_name in ("forward", "forward_4k"):
        npx = B * H * W
        stats = torch.tensor([float(torch.log2(out["likelihoods"]["y_likelihoods"].double()).sum() +
                                    torch.log2(out["likelihoods"]["z_likelihoods"].double()).sum()),
                              float(((out["x_hat"].double() - x.double()) ** 2).sum()), float(npx)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(stats)
        quality = {"bpp": -float(stats[0]) / float(stats[2]), "mse": float(stats[1]) / (3 * float(stats[2]))}

    # end to end through the public API with HOST buffers: pinned inputs -> H2D -> walk -> D2H of the results, all inside the timed region
    e2e = None
    if not args.no_e2e:
        keep = [host_step() for _ in range(2)]      # two generations of pinned result buffers into torch's cache: no cudaHostAlloc in the timed region
        del keep
        for _ in range(max(args.warmup - 2, 1)):
            o = host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            o = host_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())

        def nbytes(v):
            if torch.is_tensor(v):
                return v.numel() * v.element_size() if v.device.type == "cpu" else 0
            if isinstance(v, dict):
                return sum(nbytes(u) for u in v.values())
            if isinstance(v, (list, tuple)):
                return sum(nbytes(u) for u in v)
            return 0
        e2e = {"value": units_per_step * args.steps / dt, "unit": "MP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": nbytes(o)}

    if rank == 0:
        pk = peaks()
        tc_tflops = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        top_tflops = top_flops * top_n / (top_ms * 1e-3) / 1e12 if top_ms > 0 else 0.0
        whole = (world * B * args.steps * cfg["flop"] / (ms * 1e-3) / 1e12 / world) if (cfg["flop"] and not bands) else \
                (args.steps * cfg["flop"] / (ms * 1e-3) / 1e12 / world if cfg["flop"] else None)
        # DRAM traffic of the heaviest launch shape: one `ncu --set full` capture of conv3_pair_kernel (profiles/r02_ncu_conv3_pair.txt), taken
        # at one image per launch (dram__bytes_read.sum + dram__bytes_write.sum = 52.8 + 142.8 MB; algorithmic 50.1 + 200.5 MB, the tail of
        # the output is still in L2 when the kernel ends); scaled to this run's images per launch.  Only valid for the MLICPP_L forward.
        traffic = 195.7e6 * (4 if H == 2176 else 1) * B if cfg_name in ("forward", "forward_4k") else None
        roof = {"bound": "tensor",
                # the numbers a reader should look at first: every tensor-core launch of the step, and the whole step
                "all_tcgen05_launches": {"achieved": tc_tflops, "unit": "TFLOP/s", "frac_of_sustained": tc_tflops / pk["tf_sust"],
                                         "frac_of_burst": tc_tflops / pk["tf_burst"], "launches": tc_launches,
                                         "ms_per_step": tc_ms / max(args.steps, 1), "share_of_step": tc_ms / ms_prof if ms_prof > 0 else None},
                "whole_step": ({"achieved": whole, "unit": "TFLOP/s (algorithmic FLOP of SURVEY.md 8d / step time, per GPU)",
                                "frac_of_sustained": whole / pk["tf_sust"], "frac_of_burst": whole / pk["tf_burst"]} if whole else None),
                # the contract's single-kernel line: the heaviest launch shape of the step
                "kernel": ("heaviest tcgen05 launch shape of the step (MLICPP_L forward: conv3_pair_kernel, the two-SM 3x3 192->768 sub-pixel conv "
                           "of g_s stage 5, 2 launches per step)" if cfg_name != "sd_decode" else
                           "heaviest tcgen05 launch shape of the step (MLICPP_M_SMALL_DEC decoder walk: an HBM-bound depthwise-separable block of g_s at full "
                           "resolution, quoted here against the tensor peak like the other configs; see DESIGN.md 4.2)"),
                # against the BURST cuBLAS figure: inside the step this kernel runs above the sustained one (1361 TFLOP/s), so the stricter
                # denominator is the honest one
                "achieved": top_tflops, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": top_tflops / pk["tf_burst"],
                "frac_of_sustained": top_tflops / pk["tf_sust"],
                "peak_source": pk["src"] + " bf16_tflops (burst, best of 10 cuBLAS 8192^3); sustained " + f"{pk['tf_sust']:.1f}",
                "flops_per_launch": top_flops, "launches": top_n, "avg_launch_ms": top_ms / max(top_n, 1),
                "traffic": traffic, "traffic_source": "ncu --set full at 1 image/launch x images per launch" if traffic else None,
                "share_of_step": top_ms / ms_prof if ms_prof > 0 else None,
                "measured_in": f"second pass of the same {args.steps} steps with per-launch CUDA events ({ms_prof / max(args.steps, 1):.2f} ms/step with events)"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            mps, times, hw = cpu_baseline_mps(cfg_name, cfg, threads)
            cpu = {"value": mps, "unit": "MP/s", "cores": threads, "kind": "port",
                   "sample": f"1 image {hw[1]}x{hw[0]} of the workload per run, best of {len(times)} runs ({sum(times):.1f} s of CPU work), "
                             f"fp32 torch CPU (oracle port of the reference, weights from the oracle/shapes template), {threads} threads"}
        line = {"metric": cfg["metric"], "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "strong" if bands else "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": cfg["workload"], "images_per_gpu_per_step": B if not bands else None,
                           "global_batch": B * world if not bands else 1,
                           "parallelism": f"row bands x{world}" if bands else f"batch-shard x{world}", "weights": "random-init seed 1234",
                           "l2": "per-step inputs + activations (>1 GB/image) exceed the 126 MB L2; no explicit flush"},
                "e2e": e2e, "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu, "quality": quality}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
