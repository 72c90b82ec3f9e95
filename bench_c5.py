#!/usr/bin/env python
"""bench_c5.py -- BASELINE.json configs[4]: the reference's own training step with the GPU alignment
drop-in, and the alignment's share of the step time.

    python bench.py --workload c5 [--gpus N] [--steps K] [--warmup W]      (bench.py forwards here)

The model code is the UNMODIFIED reference (`SynthesizerTrn`, `AvocodoDiscriminator`, `losses`, `commons`,
`monotonic_align/__init__.py` + its own compiled Cython core), installed into the git-ignored
``baseline/_ref/vits`` by ``tools/install_reference.py``.  `train.py` / `train_and_evaluate.py` themselves do
not import in this image (omegaconf, phaseaug, librosa, matplotlib are absent), so ``train_step``
below restates one iteration of the reference's loop (`train_and_evaluate.py:55-156`) around the reference's
modules: autocast(fp16_run) generator forward through DDP, discriminator step, generator step, GradScaler,
AdamW with the config's hyper-parameters (`train.py:142-176,201`).  Stated departures, none on the alignment path:
PhaseAug is the identity (`train_and_evaluate.py:95-98,118-121`), the mel L1 term (`:127`, librosa filterbank) is
an L1 on STFT magnitudes of the same segments, data are synthetic tensors of the collate's shapes
(`TextAudioSpeakerCollate.py:96`).

Arms (same seeded model + batch in every arm):
  stock          `SynthesizerTrn.maximum_path` = the reference's wrapper + Cython (`monotonic_align/__init__.py:7-20`)
  dropin         `SynthesizerTrn.maximum_path = vits_b200.maximum_path`  
  dropin_stats   as `dropin`, and the path is computed from (z_p, m_p, logs_p) by
                 `vits_b200.maximum_path_from_stats` (contraction + search in one call); the reference's inline
                 einsums (`SynthesizerTrn.py:223-232`) still execute because `forward` is not edited, so this arm's
                 `alignment_ms` is an upper bound for a maintainer who deletes those ten lines.

Timed with CUDA events on the training stream: `step_ms` = one whole iteration; `alignment_ms` = from the end of
`self.flow(...)` (`SynthesizerTrn.py:210`, a forward hook) to the return of `maximum_path` (`:235`) -- the three
`torch.cat` of `:212-215` fall inside, they are microseconds; `mas_ms` = the `maximum_path` call alone.
"""
from __future__ import annotations

import importlib
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
REF_VITS = os.path.join(ROOT, "baseline", "_ref", "vits")


def load_reference():
    """Import the unmodified reference modules from baseline/_ref/vits (their own `monotonic_align` included)."""
    if not os.path.isdir(REF_VITS):
        raise RuntimeError("baseline/_ref/vits is missing: run `python tools/install_reference.py` where /root/reference exists")
    if REF_VITS not in sys.path:
        sys.path.insert(0, REF_VITS)
    for name in ("monotonic_align",):
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF_VITS):
            del sys.modules[name]       # never let another `monotonic_align` shadow the reference's
    ST = importlib.import_module("SynthesizerTrn")
    AV = importlib.import_module("Avocodo")
    losses = importlib.import_module("losses")
    commons = importlib.import_module("commons")
    ref_mas = importlib.import_module("monotonic_align")
    assert ref_mas.__file__.startswith(REF_VITS)
    return ST, AV, losses, commons, ref_mas


def load_config():
    import yaml
    with open(os.path.join(REF_VITS, "configs", "config_cje.yaml")) as f:
        cfg = yaml.safe_load(f)
    for k in ("learning_rate", "eps", "lr_decay"):      # PyYAML reads `2e-4` as a string (OmegaConf does not)
        cfg["train"][k] = float(cfg["train"][k])
    return cfg


def build_models(ST, AV, cfg, device, seed=1234):
    import torch
    torch.manual_seed(seed)
    data, model = cfg["data"], cfg["model"]
    net_g = ST.SynthesizerTrn(                                            # train.py:142-151 (71 = len(symbols))
        71, data["filter_length"] // 2 + 1, cfg["train"]["segment_size"] // data["hop_length"],
        n_speakers=len(data["speakers"]), midi_start=data["midi_start"], midi_end=data["midi_end"],
        octave_range=data["octave_range"], **model).to(device)
    net_d = AV.AvocodoDiscriminator(model["use_spectral_norm"]).to(device)  # train.py:153
    return net_g, net_d


def synthetic_batch(cfg, B, T_x, T_y, device, seed, ragged=True):
    """Tensors of the collate's shapes (TextAudioSpeakerCollate.py:96); lengths drawn like SURVEY.md 8(d), sorted by
    spec length descending like the collate (:26-30); padded tails zero."""
    import torch
    rng = np.random.default_rng(seed)
    data = cfg["data"]
    if ragged:
        t_xs = rng.integers((T_x + 1) // 2, T_x + 1, size=B)
        t_ys = np.array([rng.integers(max(tx, (T_y + 1) // 2), T_y + 1) for tx in t_xs])
        t_xs[0], t_ys[0] = T_x, T_y
        order = np.argsort(-t_ys, kind="stable")
        t_xs, t_ys = t_xs[order], t_ys[order]
    else:
        t_xs, t_ys = np.full(B, T_x), np.full(B, T_y)
    g = torch.Generator().manual_seed(seed)
    hop = data["hop_length"]
    x = torch.randint(1, 71, (B, T_x), generator=g)
    tone = torch.randint(0, 4, (B, T_x), generator=g)
    spec = torch.rand(B, data["filter_length"] // 2 + 1, T_y, generator=g)
    ying = torch.rand(B, data["midis"] if "midis" in data else 80, T_y, generator=g)
    wav = torch.rand(B, 1, T_y * hop, generator=g) * 0.2 - 0.1
    for i in range(B):
        x[i, t_xs[i]:] = 0
        tone[i, t_xs[i]:] = 0
        spec[i, :, t_ys[i]:] = 0
        ying[i, :, t_ys[i]:] = 0
        wav[i, :, t_ys[i] * hop:] = 0
    sid = torch.randint(0, len(data["speakers"]), (B,), generator=g)
    to = lambda t: t.to(device)
    return dict(x=to(x), tone=to(tone), x_lengths=to(torch.as_tensor(t_xs)), spec=to(spec),
                spec_lengths=to(torch.as_tensor(t_ys)), ying=to(ying), y=to(wav), sid=to(sid),
                t_xs=t_xs.astype(np.int32), t_ys=t_ys.astype(np.int32))


def stft_mag(y, n_fft, hop, win):
    import torch
    w = torch.hann_window(win, device=y.device, dtype=torch.float32)
    s = torch.stft(y.float(), n_fft, hop_length=hop, win_length=win, window=w, center=True, return_complex=True)
    return torch.sqrt(s.real ** 2 + s.imag ** 2 + 1e-9)


class Taps:
    """CUDA-event taps around the alignment section without editing the reference's forward."""

    def __init__(self, ST, net_g_module, arm, ref_maximum_path):
        import torch
        self.torch = torch
        self.ST, self.arm, self.ref_mp = ST, arm, ref_maximum_path
        self.cuda = next(net_g_module.parameters()).is_cuda
        self.records = []            # per forward: (ev_align0, ev_mas0, ev_mas1)
        self.cap = {}
        self.keep = None             # last (neg_cent, mask, attn) for parity checks
        self.keep_inputs = False
        net_g_module.flow.register_forward_hook(self._after_flow)
        net_g_module.text_encoder.register_forward_hook(self._after_text)
        ST.maximum_path = self._maximum_path          # the name bound at SynthesizerTrn.py:16, called at :235

    def _ev(self):
        if not self.cuda:
            return time.perf_counter()
        e = self.torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def _after_text(self, mod, inp, out):
        self.cap["m_p"], self.cap["logs_p"] = out[1], out[2]
        self.cap["x_lengths"] = inp[2]

    def _after_flow(self, mod, inp, out):
        self.cap["z_p"] = out
        self.cap["ev0"] = self._ev()

    def _maximum_path(self, neg_cent, mask):
        e1 = self._ev()
        if self.arm == "stock":
            out = self.ref_mp(neg_cent, mask)
        elif self.arm == "dropin":
            import vits_b200
            out = vits_b200.maximum_path(neg_cent, mask)
        elif self.arm == "dropin_stats":
            import vits_b200
            y_lengths = self.cap["y_lengths"]
            out = vits_b200.maximum_path_from_stats(self.cap["z_p"], self.cap["m_p"], self.cap["logs_p"],
                                                    self.cap["x_lengths"], y_lengths).to(neg_cent.dtype)
        else:
            raise ValueError(self.arm)
        e2 = self._ev()
        self.records.append((self.cap.get("ev0"), e1, e2))
        if self.keep_inputs:
            self.keep = dict(neg_cent=neg_cent.detach().clone(), mask=mask.detach().clone(), attn=out.detach().clone(),
                             z_p=self.cap["z_p"].detach().float().clone(), m_p=self.cap["m_p"].detach().float().clone(),
                             logs_p=self.cap["logs_p"].detach().float().clone())
        return out

    def drain(self):
        """-> list of (alignment_ms, mas_ms) per recorded forward."""
        out = []
        for e0, e1, e2 in self.records:
            if self.cuda:
                out.append((e0.elapsed_time(e2), e1.elapsed_time(e2)))
            else:
                out.append(((e2 - e0) * 1e3, (e2 - e1) * 1e3))
        self.records = []
        return out


def train_step(mods, cfg, nets, optims, scaler, batch, taps, fp16):
    """One iteration of the reference's loop (train_and_evaluate.py:55-156)."""
    import torch
    import torch.nn.functional as F
    ST, AV, losses, commons, _ = mods
    net_g, net_d = nets
    optim_g, optim_d = optims
    tr, data = cfg["train"], cfg["data"]
    dev_type = "cuda" if batch["x"].is_cuda else "cpu"
    amp = lambda enabled: torch.autocast(dev_type, dtype=torch.float16 if dev_type == "cuda" else torch.bfloat16, enabled=enabled)
    taps.cap["y_lengths"] = batch["spec_lengths"]
    with amp(fp16):
        (y_hat, l_length, attn, ids_slice, x_mask, z_mask, y_hat_, (z, z_p, m_p, logs_p, m_q, logs_q), _,
         (z_spec, m_spec, logs_spec, spec_mask, z_yin, m_yin, logs_yin, yin_mask),
         (yin_gt_crop, yin_gt_shifted_crop, yin_dec_crop, yin_hat_crop, scope_shift, yin_hat_shifted)) = net_g(
            batch["x"], batch["tone"], batch["x_lengths"], batch["spec"], batch["spec_lengths"], batch["ying"],
            batch["spec_lengths"], batch["sid"])                                                      # :55-60
        seg = tr["segment_size"]
        hop = data["hop_length"]
        yin_gt_crop = commons.slice_segments(torch.cat([yin_gt_crop, yin_gt_shifted_crop], dim=0), ids_slice, seg // hop)  # :82-85
        y_ = commons.slice_segments(torch.cat([batch["y"], batch["y"]], dim=0), ids_slice * hop, seg)   # :87-91
        aug_y_hat_ = [t.detach() for t in y_hat_]                                                     # :95-100 (PhaseAug = identity)
        y_d_hat_r, y_d_hat_g, _, _ = net_d(y_, aug_y_hat_)                                            # :102
        with amp(False):
            loss_disc, _, _ = losses.discriminator_loss(y_d_hat_r, y_d_hat_g)                          # :104-108
    optim_d.zero_grad()
    scaler.scale(loss_disc).backward()                                                                # :110-114
    scaler.unscale_(optim_d)
    commons.clip_grad_value_(net_d.parameters(), None)
    scaler.step(optim_d)
    with amp(fp16):
        y_d_hat_r, y_d_hat_g, fmap_r, fmap_g = net_d(y_, y_hat_)                                      # :127
        with amp(False):
            loss_dur = torch.sum(l_length.float())
            half = y_.shape[0] // 2
            loss_mel = F.l1_loss(stft_mag(y_[:half].squeeze(1), data["filter_length"], hop, data["win_length"]),
                                 stft_mag(y_hat[-1].squeeze(1), data["filter_length"], hop, data["win_length"])) * tr["c_mel"]
            loss_kl = losses.kl_loss(z_p, logs_q, m_p, logs_p, z_mask) * tr["c_kl"]                    # :133-135
            loss_yin_dec = F.l1_loss(yin_gt_shifted_crop, yin_dec_crop) * tr["c_yin"]
            loss_yin_shift = (F.l1_loss(torch.exp(-yin_gt_crop), torch.exp(-yin_hat_crop)) * tr["c_yin"] +
                              F.l1_loss(torch.exp(-yin_hat_shifted), torch.exp(-(torch.chunk(yin_hat_crop, 2, dim=0)[1]))) * tr["c_yin"])
            loss_fm = losses.feature_loss(fmap_r, fmap_g)
            loss_gen, _ = losses.generator_loss(y_d_hat_g)
            loss_gen_all = loss_gen + loss_fm + loss_mel + loss_dur + loss_kl + loss_yin_shift + loss_yin_dec  # :152
    optim_g.zero_grad()
    scaler.scale(loss_gen_all).backward()                                                             # :154-159
    scaler.unscale_(optim_g)
    commons.clip_grad_value_(net_g.parameters(), None)
    scaler.step(optim_g)
    scaler.update()
    return loss_gen_all.detach(), loss_disc.detach(), attn


def run_arm(arm, mods, cfg, dev, rank, world, B, T_x, T_y, steps, warmup, fp16=True):
    import torch
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    ST, AV, losses, commons, ref_mas = mods
    net_g, net_d = build_models(ST, AV, cfg, dev)
    tr = cfg["train"]
    optim_g = torch.optim.AdamW(net_g.parameters(), tr["learning_rate"], betas=tr["betas"], eps=tr["eps"])  # train.py:161-173
    optim_d = torch.optim.AdamW(net_d.parameters(), tr["learning_rate"], betas=tr["betas"], eps=tr["eps"])
    taps = Taps(ST, net_g, arm, ref_mas.maximum_path)
    if dist.is_initialized():
        ids = [dev.index] if dev.type == "cuda" else None
        net_g, net_d = DDP(net_g, device_ids=ids), DDP(net_d, device_ids=ids)                              # train.py:175-176
    scaler = torch.amp.GradScaler(dev.type, enabled=fp16 and dev.type == "cuda")                            # train.py:201
    net_g.train()
    net_d.train()
    batches = [synthetic_batch(cfg, B, T_x, T_y, dev, 1234 + 17 * rank + i) for i in range(2)]
    sync = (lambda: torch.cuda.synchronize()) if dev.type == "cuda" else (lambda: None)
    status = []

    def one(i):
        out = train_step(mods, cfg, (net_g, net_d), (optim_g, optim_d), scaler, batches[i % 2], taps, fp16)
        if arm != "stock":
            import vits_b200
            status.append(vits_b200.status_nosync())     # the asynchronous status mirror: no synchronisation
        return out
    for i in range(warmup):
        one(i)
    sync()
    if dist.is_initialized():
        dist.barrier()
    taps.drain()
    step_ms = []
    for i in range(steps):
        t0 = taps._ev()
        one(i)
        t1 = taps._ev()
        sync()
        step_ms.append(t0.elapsed_time(t1) if dev.type == "cuda" else (t1 - t0) * 1e3)
    recs = taps.drain()
    final_status = 0
    if arm != "stock" and dev.type == "cuda":
        import vits_b200
        final_status = vits_b200.last_status(dev)
    res = dict(step_ms=float(np.median(step_ms)), alignment_ms=float(np.median([r[0] for r in recs])),
               mas_ms=float(np.median([r[1] for r in recs])), steps=steps,
               status_bits_seen=int(np.bitwise_or.reduce(np.asarray(status + [final_status], dtype=np.int64))) if arm != "stock" else None)
    res["share"] = res["alignment_ms"] / res["step_ms"]
    if dist.is_initialized():
        t = torch.tensor([res["step_ms"], res["alignment_ms"], res["mas_ms"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["step_ms"], res["alignment_ms"], res["mas_ms"] = (float(v) for v in t.tolist())
        res["share"] = res["alignment_ms"] / res["step_ms"]
    del net_g, net_d, optim_g, optim_d
    return res


def main(args, print_line=None):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "config 5 needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    else:                                   # DDP wrappers as in train.py:175-176 even on one GPU
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(29500 + os.getpid() % 2000))
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    torch.backends.cudnn.benchmark = True   # train.py:25
    mods = load_reference()
    cfg = load_config()
    B, T_y, T_x = args.c5_batch, 1024, 192
    steps, warmup = max(1, min(args.steps, 10)), max(2, min(args.warmup, 3))
    arms = {}
    if getattr(args, "impl", "ours") == "reference":          # the reference arm of c5 is its stock step
        s = run_arm("stock", mods, cfg, dev, rank, world, B, T_x, T_y, steps, warmup)
        if rank == 0:
            (print_line or (lambda l: print(json.dumps(l), flush=True)))({
                "impl": "reference", "metric": "alignment share of the SynthesizerTrn training step", "value": s["share"],
                "unit": "fraction of step time", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": s["step_ms"],
                "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "data": "synthetic", "arms": {"stock": s},
                "config": {"workload": f"c5: B={B}/GPU T_y<={T_y} T_x<={T_x}"}, "gpu_launches": 0})
        dist.destroy_process_group()
        return
    for arm in ("stock", "dropin", "dropin_stats"):
        arms[arm] = run_arm(arm, mods, cfg, dev, rank, world, B, T_x, T_y, steps, warmup)
        torch.cuda.empty_cache()
    if rank == 0:
        s, d = arms["stock"], arms["dropin"]
        line = {
            "metric": "alignment share of the SynthesizerTrn training step", "value": d["share"], "unit": "fraction of step time",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": d["step_ms"], "higher_is_better": False,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 alignment inside an autocast(fp16) step", "data": "synthetic",
            "config": {"workload": f"c5: reference SynthesizerTrn + AvocodoDiscriminator training step (train_and_evaluate.py:55-156), "
                                   f"DDP over {world} GPU(s), B={B}/GPU T_y<={T_y} T_x<={T_x} variable lengths, autocast fp16 + GradScaler",
                       "departures": "PhaseAug = identity, mel L1 -> STFT-magnitude L1 (phaseaug / librosa absent); synthetic batch"},
            "arms": arms,
            "alignment_ms": {"stock": s["alignment_ms"], "dropin": d["alignment_ms"], "dropin_stats": arms["dropin_stats"]["alignment_ms"]},
            "step_ms": {"stock": s["step_ms"], "dropin": d["step_ms"], "dropin_stats": arms["dropin_stats"]["step_ms"]},
            "share": {"stock": s["share"], "dropin": d["share"], "dropin_stats": arms["dropin_stats"]["share"]},
            "step_speedup_vs_stock": s["step_ms"] / d["step_ms"],
            "timeout_status_seen": bool((d["status_bits_seen"] or 0) & 8 or (arms["dropin_stats"]["status_bits_seen"] or 0) & 8),
            "gpu_launches": None,
        }
        from vits_b200 import _lib
        line["gpu_launches"] = _lib.launch_count()
        (print_line or (lambda l: print(json.dumps(l), flush=True)))(line)
    dist.destroy_process_group()
