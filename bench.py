#!/usr/bin/env python
"""bench.py -- SDS train-steps/sec (NeRF render + SD U-Net) @128x128 on N B200s (BASELINE.json metric).

One step == the SDS branch of Trainer.train_step (nerf/utils.py:461-574) restricted to the hot path:
    model.render(128x128 rays, perturb, force_all_rays, max_steps 512, k = 13 field evaluations / sample)
    -> guidance.train_step(text_z, pred_rgb)   [bilinear 512 -> VAE encode -> add_noise -> U-Net x2 (CFG) -> SDS grad
                                                -> latents.backward: VAE input-grad -> composite bwd -> field bwd]
    -> regulariser loss.backward()             [opacity + entropy + orientation + smoothness: second render backward]
    -> (N > 1) all-reduce of hash-grid + MLP gradients
Excluded like SURVEY 8d says: CLIP losses, PNG dumps, the Adan update.  Synthetic data: orbit poses, solid-sphere occupancy
(r = 0.2), seeded random hash table / MLP / SD-2.0-base weights (no weights offline), N(0,1) text embeddings.

  python bench.py --gpus 1 --steps 10 --warmup 3            # one JSON line
  torchrun ... bench.py --gpus N ...                        # one rank per GPU, weak scaling (one view per GPU per step)
  python bench.py --impl reference ...                      # the reference algorithm on the host CPU cores (oracle port)
"""
import argparse
import importlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SDS train-steps/sec (NeRF render+SD U-Net) @128x128"
UNIT = "view-steps/s"
HW = 128
SPHERE_R = 0.2
WORKLOAD = ("coarse stage 128x128, SD-2.0-base architecture (random-init weights) SDS, hash-grid L=16 F=2 T=2^19 fp32, "
            "solid-sphere occupancy r=0.2, max_steps 512, k=13 field evals/sample (reference defaults lambda_orient/lambda_smooth on)")


def opt_namespace():
    return argparse.Namespace(bound=1, min_near=0.1, density_thresh=10, bg_radius=-1, blob_density=5, blob_radius=0.1,
                              lambda_smooth=1, lambda_orient=1e-2, lambda_opacity=1e-3, lambda_entropy=1, max_depth=10.0,
                              max_steps=512, dt_gamma=0)


def orbit_pose(radius, theta_deg, phi_deg):
    """look-at camera of nerf/provider.py:143-214 (up = -y)"""
    th, ph = math.radians(theta_deg), math.radians(phi_deg)
    c = np.array([radius * math.sin(th) * math.sin(ph), radius * math.cos(th), radius * math.sin(th) * math.cos(ph)], np.float32)
    fwd = -c / np.linalg.norm(c)
    right = np.cross(fwd, np.array([0, -1, 0], np.float32)); right /= np.linalg.norm(right)
    up = np.cross(right, fwd); up /= np.linalg.norm(up)
    pose = np.eye(4, dtype=np.float32)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, up, fwd, c
    return pose


def synth_pose(index, rng):
    """radius U[1,1.5], theta U[70,110], phi U[0,360), fovy U[15,25] (main.py:72-76); index % 4 == 0 is the front view."""
    if index % 4 == 0:
        return orbit_pose(1.0, 90.0, 180.0), 20.0
    return orbit_pose(rng.uniform(1.0, 1.5), rng.uniform(70, 110), rng.uniform(0, 360)), rng.uniform(15, 25)


def sphere_bitfield_numpy(radius, H=128):
    def spread(v):
        v = (v * 0x00010001) & 0xFF0000FF; v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3; v = (v * 0x00000005) & 0x49249249
        return v
    idx = np.arange(H, dtype=np.uint32)
    s = spread(idx.astype(np.uint64)).astype(np.uint32)
    xs, ys, zs = np.meshgrid(idx, idx, idx, indexing="ij")
    morton = (s[xs] | (s[ys] << 1) | (s[zs] << 2)).reshape(-1)
    c = (np.stack([xs, ys, zs], -1).reshape(-1, 3).astype(np.float32) + 0.5) / H * 2 - 1
    occ = np.zeros(H ** 3, np.uint8)
    occ[morton] = (np.linalg.norm(c, axis=1) < radius)
    return np.packbits(occ, bitorder="little")


class ClockSampler:
    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_models(device, seed):
    nt = importlib.import_module("make-it-3d_b200.nerf.network_tcnn")
    sdm = importlib.import_module("make-it-3d_b200.nerf.sd")
    torch.manual_seed(seed)
    model = nt.NeRFNetwork(opt_namespace())
    with torch.no_grad():
        g = torch.Generator().manual_seed(3)
        # "trained-like" table amplitude so densities / colours vary (init U(-1e-4,1e-4) would make the field a pure blob)
        model.encoder.params.copy_((torch.rand(model.encoder.params.numel(), generator=g) * 2 - 1) * 0.5)
    model = model.to(device).train()
    model.density_bitfield = torch.from_numpy(sphere_bitfield_numpy(SPHERE_R)).to(device)
    guidance = sdm.StableDiffusion(device, seed=0)
    return model, guidance


def regulariser_loss(out, opt):
    """nerf/utils.py:519-548 for global_step >= diff_iters (opacity, entropy x10, orientation x(1+10), smoothness)."""
    ws = out["weights_sum"]
    loss = opt.lambda_opacity * (ws ** 2).mean()
    a = ws.clamp(1e-5, 1 - 1e-5)
    loss = loss + opt.lambda_entropy * 10 * (-a * torch.log2(a) - (1 - a) * torch.log2(1 - a)).mean()
    loss = loss + opt.lambda_orient * out["loss_orient"] * 11
    loss = loss + opt.lambda_smooth * out["loss_smooth"]
    return loss


def run_ours(args):
    par = importlib.import_module("make-it-3d_b200.parallel")
    field_ops = importlib.import_module("make-it-3d_b200.nerf.field_ops")
    utils = importlib.import_module("make-it-3d_b200.nerf.utils")
    rank, local, world = par.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    model, guidance = build_models(device, seed=0)
    opt = model.opt
    reducer = par.GradientAllReduce(model.encoder.params, list(model.sigma_net.parameters()), op="sum")
    rng = np.random.default_rng(par.rank_seed(0, rank))
    torch.manual_seed(par.rank_seed(0, rank))
    text_z = torch.randn(2, 77, 1024, generator=torch.Generator().manual_seed(0)).to(device)
    n_steps = args.warmup + args.steps + 4
    # ---- host-side (pinned) inputs for the end-to-end leg; device copies for the device-resident leg ----
    host_in, dev_in = [], []
    for s in range(n_steps):
        pose, fov = synth_pose(par.pose_index(s, rank, world), rng)
        focal = HW / (2 * math.tan(math.radians(fov) / 2))
        rays = utils.get_rays(torch.from_numpy(pose)[None], (focal, focal, HW / 2, HW / 2), HW, HW, -1)
        packed = torch.cat([rays["rays_o"].reshape(-1), rays["rays_d"].reshape(-1), rays["depth_scale"].reshape(-1)]).contiguous().pin_memory()
        host_in.append(packed)
        dev_in.append(packed.to(device))
    text_host = text_z.cpu().pin_memory()
    N = HW * HW
    h2d_bytes = host_in[0].numel() * 4 + text_host.numel() * 4
    t_cycle = (250, 450, 600)          # islarge=True keeps every step on the SDS branch (nerf/sd.py:153)

    def unpack(buf):
        return buf[:3 * N].view(1, N, 3), buf[3 * N:6 * N].view(1, N, 3), buf[6 * N:].view(1, N)

    result_dev = torch.zeros(4, device=device)
    result_host = torch.zeros(4).pin_memory()

    def step(s, e2e):
        if e2e:
            buf = host_in[s % n_steps].to(device, non_blocking=True)
            tz = text_host.to(device, non_blocking=True)
        else:
            buf, tz = dev_in[s % n_steps], text_z
        rays_o, rays_d, depth_scale = unpack(buf)
        model.zero_grad(set_to_none=True)
        bg = torch.rand(3, device=device)
        out = model.render(rays_o, rays_d, depth_scale=depth_scale, bg_color=bg, staged=False, perturb=True, ambient_ratio=1.0,
                           shading='albedo', force_all_rays=True, **vars(opt))
        pred_rgb = out['image'].reshape(1, HW, HW, 3).permute(0, 3, 1, 2).contiguous()
        loss, _ = guidance.train_step(tz, pred_rgb, islarge=True, guidance_scale=10, t=t_cycle[s % 3])
        loss = loss + regulariser_loss(out, opt)
        loss.backward()
        reducer()
        if e2e:
            result_dev[0] = loss.detach(); result_dev[1] = out['loss_orient'].detach(); result_dev[2] = out['loss_smooth'].detach()
            result_dev[3] = out['weights_sum'].mean().detach()
            result_host.copy_(result_dev, non_blocking=True)
        return loss

    def timed(n, e2e, first):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(n):
            step(first + s, e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        return par.max_over_ranks(e0.elapsed_time(e1), device)

    for s in range(max(args.warmup, 3)):
        step(s, False)
    torch.cuda.synchronize()
    if args.launch_list:
        torch.cuda.nvtx.range_push("timed_steps")
        for s in range(args.steps):
            step(args.warmup + s, False)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    # ---- launch count (one profiled step, outside the timed region) ----
    launches_per_step = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            step(0, False)
            torch.cuda.synchronize()
        mine = ("k_field", "k_bwd", "k_march", "k_composite", "k_tc_gemm", "sdk::", "sd::k_", "k_loss_finalize", "k_near_far", "k_packbits", "k_grid")
        launches_per_step = sum(1 for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and any(m in ev.name for m in mine))
    except Exception:
        launches_per_step = None
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    ms = timed(args.steps, False, args.warmup)
    # ---- per-kernel live timing of the dominant kernels (CUDA events on the launching stream) ----
    import ctypes as C
    L = importlib.import_module("make-it-3d_b200._lib")
    field_ops.PROFILE = []
    L.check(L.lib().mi3d_sd_profile(guidance.engine.h, C.c_int(1), None, None), "sd_profile")
    n_prof = 3
    for s in range(n_prof):
        step(args.warmup + s, False)
    torch.cuda.synchronize()
    gemm_ms, gemm_n = C.c_float(0), C.c_int(0)
    dump_path = os.path.join(ROOT, "gpurun_out", f"sd_launches_rank{rank}.txt")
    os.makedirs(os.path.dirname(dump_path), exist_ok=True)
    os.environ["MI3D_SD_PROFILE_DUMP"] = dump_path          # one line per timed launch: M N K block_n splits conv batch epi ms
    L.check(L.lib().mi3d_sd_profile(guidance.engine.h, C.c_int(0), C.byref(gemm_ms), C.byref(gemm_n)), "sd_profile")
    tile_flops = attn_flops = attn_ms = 0.0
    attn_n = 0
    for ln in open(dump_path):
        f = ln.split()
        Mm, Nn, Kk, conv, batch, ms_l = int(f[0]), int(f[1]), int(f[2]), int(f[5]), int(f[6]), float(f[8])
        if conv == 2:
            attn_flops += 4.0 * Mm * Nn * Kk * batch; attn_ms += ms_l; attn_n += 1
        else:
            tile_flops += 2.0 * Mm * Nn * Kk * batch
    prof_rows = field_ops.PROFILE
    field_ops.PROFILE = None
    for s in range(2):                      # untimed: first-use allocations of the host-input staging path
        step(s, True)
    ms_e2e = timed(args.steps, True, args.warmup)
    if os.environ.get("MI3D_BENCH_ABAB"):        # diagnostic: is the e2e/device gap the copies or the order (clocks)?
        a2 = timed(args.steps, False, args.warmup); b2 = timed(args.steps, True, args.warmup); a3 = timed(args.steps, False, args.warmup)
        print(f"[abab] dev {ms / args.steps:.3f} e2e {ms_e2e / args.steps:.3f} dev {a2 / args.steps:.3f} e2e {b2 / args.steps:.3f} dev {a3 / args.steps:.3f} ms/step", file=sys.stderr)
    clock_info = clocks.stop()
    ws = list(model._workspaces.values())[0]
    M = int(ws.counter[0])

    value = world * args.steps / (ms * 1e-3)
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32 render / fp16 SD (fp32 accumulate)",
        "data": "synthetic (orbit poses, sphere occupancy, random-init hash table / MLP / SD-2.0-base weights, N(0,1) text embeddings)",
        "config": {"workload": WORKLOAD, "rays": N, "samples_per_step_M": M, "field_evals_per_sample": 13, "views_per_step": world,
                   "parallelism": f"view-dp{world}", "l2": "inputs larger than L2: 1.8 GB of SD weights + activations stream through L2 every step "
                   "(the 48.8 MB hash table is re-fetched after each SD pass)", "excluded": "CLIP losses, PNG I/O, Adan update (SURVEY 8d)"},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(result_host.numel() * 4)},
        "gpu_launches": (launches_per_step * args.steps) if launches_per_step is not None else None,
        "clocks": clock_info,
    }
    # roofline of the dominant kernel = the tcgen05 tile kernel k_tc_gemm (every conv / linear / attention product of the U-Net
    # and the VAE: the largest share of the step).  achieved = algorithmic FLOPs of one step (1.608 + 1.117 + 1.117
    # TFLOP) / summed duration of its launches in that step, timed live with CUDA events around each launch.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    per_kernel = {}
    for name, s, e, info in prof_rows:
        key = name + ("_full" if info.get("full") else ("_image" if name == "k_field_bwd" else ""))
        per_kernel.setdefault(key, []).append(s.elapsed_time(e))
    m_pad = M + 128 - M % 128
    summary = {}
    try:
        summary = json.load(open(os.path.join(ROOT, "profiles", "r1_summary.json")))
    except Exception:
        pass
    gemm_ms_step = gemm_ms.value / n_prof
    flops = tile_flops / n_prof          # 2 M N K of every tile-kernel launch of one step (the engine's own launch list)
    line["roofline"] = {
        "kernel": "tc::k_tc_gemm<64|128|160|256> (tcgen05.mma/TMEM/TMA tile kernel; all launches of one step: every conv / linear of the U-Net and the VAE)", "bound": "tensor",
        "achieved": round(flops / (gemm_ms_step * 1e-3) / 1e12, 1), "peak": tf_peak, "unit": "TFLOP/s",
        "frac": round(flops / (gemm_ms_step * 1e-3) / 1e12 / tf_peak, 4),
        "traffic": summary.get("k_tc_gemm", {}).get("dram_bytes_per_launch"),
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured; kernel timed inside a long step)" if peaks else "fallback 1400 TFLOP/s (of fallback)",
        "ms_per_step": round(gemm_ms_step, 3), "launches_per_step": gemm_n.value // n_prof, "algorithmic_flops_per_step": flops,
    }
    if attn_n:
        line["roofline"]["attention"] = {"kernel": "attn::k_flash_attn (tcgen05, scores in TMEM)", "ms_per_step": round(attn_ms / n_prof, 3),
                                         "launches_per_step": attn_n // n_prof, "achieved": round(attn_flops / (attn_ms * 1e-3) / 1e12, 1), "unit": "TFLOP/s",
                                         "frac": round(attn_flops / (attn_ms * 1e-3) / 1e12 / tf_peak, 4)}
    # second roofline: the render kernels against the HBM roofline the north_star names (algorithmic bytes of SURVEY 8d;
    # these kernels are really bound by L1 gather / RED-atomic throughput)
    rk = {k: round(float(np.median(v)), 3) for k, v in per_kernel.items()}
    fwd_ms = rk.get("k_field_fwd")
    bwd_ms = rk.get("k_field_bwd_full")
    line["roofline_render"] = {"bound": "hbm", "peak": hbm, "unit": "GB/s", "kernels_ms": rk,
                               "traffic": {k: summary.get(k, {}).get("dram_bytes_per_launch") for k in ("k_field_fwd_tc", "k_bwd_enc_scatter", "k_field_bwd_tc")}}
    if fwd_ms:
        ab = 13 * m_pad * 1024 + N * 44 + 262144
        line["roofline_render"]["fwd"] = {"algorithmic_bytes": int(ab), "achieved": round(ab / (fwd_ms * 1e-3) / 1e9, 1), "frac": round(ab / (fwd_ms * 1e-3) / 1e9 / hbm, 4)}
    if bwd_ms:
        ab = 13 * m_pad * 2048 + N * 16
        line["roofline_render"]["bwd"] = {"algorithmic_bytes": int(ab), "achieved": round(ab / (bwd_ms * 1e-3) / 1e9, 1), "frac": round(ab / (bwd_ms * 1e-3) / 1e9 / hbm, 4)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def cpu_baseline(args, steps=1):
    """The oracle port of the reference algorithm (oracle/: C ray-march + PyTorch fp32 field / U-Net / VAE) on the host cores.
    Bounded sample: the FULL-size SD guidance step once + the render forward/backward on a 1/64 ray subset (16x16 rays spanning the
    128x128 view's field of view, k = 13), extrapolated linearly in the ray count.  A reported baseline, not the target."""
    from oracle import field_ref as fr
    from oracle import sd_ref
    # more threads than ~32 make torch's CPU kernels slower on these many-core hosts (measured: 128 threads -> 30x slower)
    torch.set_num_threads(min(32, os.cpu_count()))
    cores = torch.get_num_threads()
    sub = 16
    pose = orbit_pose(1.25, 80.0, 170.0)
    focal = sub / (2 * math.tan(math.radians(20.0) / 2))
    ro, rd, sc = fr.get_rays_ref(pose, (focal, focal, sub / 2, sub / 2), sub, sub)
    field = fr.FieldRef(seed=0, table_scale=0.5)
    bits = sphere_bitfield_numpy(SPHERE_R)
    rng = np.random.default_rng(0)
    t0 = time.time()
    out = fr.render_train_ref(field, ro, rd, bits, noises=rng.random(sub * sub, dtype=np.float32), light_d=np.array([0, 0.6, 0.8], np.float32),
                              bg_color=rng.random(3, dtype=np.float32), depth_scale=sc, max_steps=512, shading="albedo", lambda_smooth=1.0)
    (out["image"].sum() + out["loss_orient"] + out["loss_smooth"]).backward()
    t_render = (time.time() - t0) * (HW * HW) / (sub * sub)
    unet = sd_ref.UNet2DConditionModel(sd_ref.sd20_unet_config()).eval()
    vae = sd_ref.AutoencoderKLEncoder(sd_ref.sd_vae_config()).eval()
    for p in list(unet.parameters()) + list(vae.parameters()):
        p.requires_grad_(False)
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(1, 3, HW, HW, generator=g).requires_grad_()
    t0 = time.time()
    sd_ref.sds_train_step_ref(unet, vae, torch.randn(2, 77, 1024, generator=g), rgb, 500, torch.randn(1, 4, 64, 64, generator=g),
                              torch.randn(1, 4, 64, 64, generator=g), guidance_scale=10.0)
    t_sd = time.time() - t0
    return {"value": round(1.0 / (t_render + t_sd), 5), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"1 full-size SD guidance step ({t_sd:.1f} s) + render fwd/bwd on a 16x16 ray subset extrapolated x64 ({t_render:.1f} s extrapolated)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
            "ms_per_step": round(1000.0 / cb["value"], 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "note": "reference algorithm (oracle port: C ray-march + PyTorch fp32 field/U-Net/VAE) on host CPU cores; "
                                            "the reference's own CUDA/tcnn/diffusers path cannot run without its third-party packages and SD weights"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--launch-list", action="store_true",
                    help="profiling aid: warm up, run --steps steps of the same step function and exit (use under "
                         "`ncu --metrics gpu__time_duration.sum`; numbers printed under a profiler are never bench values)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
