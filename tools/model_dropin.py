"""Whole-model drop-in run: the UNMODIFIED reference model on top of the B200 extension.

BASELINE.json configs[2] (ResNet-50 inference on a 36-frame 360x640 clip, one expression) and
configs[3] (Video-Swin-T training step, 2 clips x 5 frames per GPU, DDP) need the reference's own
Python model (models/tce_rvos.py, models/tce_deformable_transformer.py, engine-side criterion ...).
That code is out of scope to rewrite and must run untouched on the new operator.

How it is made to run on the GPU box, where /root/reference does not exist:

  python tools/model_dropin.py --stage        (build container) copies the reference's PYTHON files
                                              into baseline/_ref/tce_rvos_src/ — git-ignored, never
                                              committed, shipped to the box by gpurun like the other
                                              contents of baseline/_ref/.  No file is edited.
  python tools/model_dropin.py --config 3|4   runs on the box.

Everything the reference needs that this image lacks is supplied from OUTSIDE its files: import
shims (tools/shims), a random-init RoBERTa + whitespace tokenizer patched onto
`transformers.*.from_pretrained` (no network), `is_main_process -> False` so the backbone does not
download ImageNet weights, and `models/ops` resolving `MultiScaleDeformableAttention_update` to
  --op b200        the extension of this repo (tce-rvos_b200/_native), or
  --op reference   the reference's own CUDA kernels compiled for sm_100 (tools/build_reference_cuda.py),
so the same model, weights and inputs are timed on both operators.

Developer / evidence tool: not imported by tests/, bench.py or __graft_entry__.py.
Under torchrun it wraps the model in DistributedDataParallel exactly as main.py:67-72 does.
"""
from __future__ import annotations

import argparse
import importlib
import importlib.util
import json
import os
import shutil
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "tce_rvos_src")
REF = os.environ.get("TCE_RVOS_REF", "/root/reference")


def stage() -> int:
    if not os.path.isdir(REF):
        print(f"{REF} not found: staging is only possible in the build container")
        return 1
    if os.path.isdir(STAGED):
        shutil.rmtree(STAGED)
    keep_ext = {".py"}

    def ignore(d, names):
        out = []
        for n in names:
            p = os.path.join(d, n)
            if os.path.isdir(p):
                if n in (".git", "docs", "__pycache__", "src"):        # models/ops/src: native code, not needed
                    out.append(n)
            elif os.path.splitext(n)[1] not in keep_ext:
                out.append(n)
        return out

    shutil.copytree(REF, STAGED, ignore=ignore)
    n = sum(len(f) for _, _, f in os.walk(STAGED))
    print(f"staged {n} python files under {STAGED} (git-ignored)")
    return 0


# ------------------------------------------------------------------------------------------------
# environment the reference expects, supplied from outside its files
# ------------------------------------------------------------------------------------------------
def swap_module_class():
    """--module b200: the reference's transformer imports `MSDeformAttn` from models.ops.modules
    (models/tce_deformable_transformer.py:20); hand it this repo's mirror of the class instead (same
    constructor, state_dict keys and 3-tuple return), again without editing a reference file."""
    pkg = importlib.import_module("tce-rvos_b200")
    ref_modules = importlib.import_module("models.ops.modules")
    ref_modules.MSDeformAttn = pkg.MSDeformAttn
    importlib.import_module("models.ops.modules.ms_deform_attn").MSDeformAttn = pkg.MSDeformAttn
    # `import models` has already run models/tce_deformable_transformer.py:20, which bound the name
    importlib.import_module("models.tce_deformable_transformer").MSDeformAttn = pkg.MSDeformAttn
    return pkg.MSDeformAttn


def install_environment(op: str, cpu_stub: bool):
    import torch

    src = STAGED if os.path.isdir(STAGED) else REF
    if not os.path.isdir(src):
        raise SystemExit("reference python tree not found: run `python tools/model_dropin.py --stage` in the build container")
    sys.path.insert(0, os.path.join(ROOT, "tools", "shims"))
    sys.path.insert(0, src)
    sys.path.insert(0, ROOT)

    if cpu_stub:
        # CPU dry run of the harness only (no GPU in the build container): the reference's OWN
        # pure-PyTorch path stands in for the extension.  Never used for a measurement.
        import types
        stub = types.ModuleType("MultiScaleDeformableAttention_update")
        sys.modules["MultiScaleDeformableAttention_update"] = stub
        fn = importlib.import_module("models.ops.functions.ms_deform_attn_func")

        def fwd(value, shapes, lsi, loc, aw, step, is_3d):
            return fn.ms_deform_attn_core_pytorch(value, shapes, loc, aw)

        def bwd(value, shapes, lsi, loc, aw, grad_out, step, is_3d):
            with torch.enable_grad():
                v, l, w = (t.detach().requires_grad_(True) for t in (value, loc, aw))
                out = fn.ms_deform_attn_core_pytorch(v, shapes, l, w)
                return list(torch.autograd.grad(out, (v, l, w), grad_out))

        stub.ms_deform_attn_forward = fwd
        stub.ms_deform_attn_backward = bwd
        ext_name = "cpu-stub(reference ms_deform_attn_core_pytorch)"
    elif op == "b200":
        pkg = importlib.import_module("tce-rvos_b200")
        pkg.install_drop_in()
        ext = pkg.load_extension()
        ext_name = ext.__file__
    else:
        path = os.path.join(ROOT, "baseline", "_ref", "MSDA_reference_cuda.so")
        spec = importlib.util.spec_from_file_location("MSDA_reference_cuda", path)
        ext = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ext)
        sys.modules["MultiScaleDeformableAttention_update"] = ext
        ext_name = path

    # text encoder: random-init RoBERTa-base, whitespace tokenizer (no network, no caches)
    import transformers
    from transformers import RobertaConfig, RobertaModel
    from transformers.tokenization_utils_base import BatchEncoding

    class _Tokenizer:
        def batch_encode_plus(self, captions, padding="longest", return_tensors="pt"):
            ids = [[0] + [3 + (zlib.crc32(w.encode()) % 50000) for w in c.lower().split()] + [2] for c in captions]
            width = max(len(i) for i in ids)
            input_ids = torch.full((len(ids), width), 1, dtype=torch.long)
            mask = torch.zeros((len(ids), width), dtype=torch.long)
            for r, i in enumerate(ids):
                input_ids[r, :len(i)] = torch.tensor(i)
                mask[r, :len(i)] = 1
            return BatchEncoding({"input_ids": input_ids, "attention_mask": mask})

    transformers.RobertaTokenizerFast.from_pretrained = classmethod(lambda cls, *a, **k: _Tokenizer())
    transformers.RobertaModel.from_pretrained = classmethod(
        lambda cls, *a, **k: RobertaModel(RobertaConfig(vocab_size=50265, max_position_embeddings=514,
                                                        type_vocab_size=1, pad_token_id=1)))
    import models.backbone as backbone_mod
    backbone_mod.is_main_process = lambda: False          # models/backbone.py:94-96: no ImageNet download
    return ext_name


def reference_args(extra):
    import opts
    parser = argparse.ArgumentParser(parents=[opts.get_args_parser()])
    args = parser.parse_args(extra)
    # main.py / inference scripts read these although opts.py never defines them
    for k, v in (("cyclic_lr", False), ("cyclic_lr_boundary", None), ("f_extra", 0)):
        if not hasattr(args, k):
            setattr(args, k, v)
    args.masks = True                                     # main.py:31 / inference_ytvos.py:47 force it
    return args


def count_op_calls():
    """Wrap the extension entry points to count calls (evidence that the model really runs on them)."""
    ext = sys.modules["MultiScaleDeformableAttention_update"]
    import torch
    calls = {"fwd": 0, "bwd": 0, "events": {"fwd": [], "bwd": []}}
    f0, b0 = ext.ms_deform_attn_forward, getattr(ext, "ms_deform_attn_backward", None)
    timed = torch.cuda.is_available()

    def run(kind, f, a):
        calls[kind] += 1
        if not timed or not a[0].is_cuda:
            return f(*a)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f(*a)
        e1.record()
        calls["events"][kind].append((e0, e1))
        return r

    def fwd(*a):
        return run("fwd", f0, a)

    def bwd(*a):
        return run("bwd", b0, a)

    fn = importlib.import_module("models.ops.functions.ms_deform_attn_func")

    class _Counted:
        ms_deform_attn_forward = staticmethod(fwd)
        ms_deform_attn_backward = staticmethod(bwd)

    fn.MSDA = _Counted
    return calls


def synthetic_targets(n_clips, T, H, W, device, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    targets = []
    for _ in range(n_clips):
        cx, cy = 0.3 + 0.4 * torch.rand(T, generator=g), 0.3 + 0.4 * torch.rand(T, generator=g)
        w, h = 0.1 + 0.2 * torch.rand(T, generator=g), 0.1 + 0.2 * torch.rand(T, generator=g)
        boxes = torch.stack([cx, cy, w, h], -1)
        masks = torch.zeros(T, H, W, dtype=torch.bool)
        for t in range(T):
            x0, x1 = int((cx[t] - w[t] / 2) * W), int((cx[t] + w[t] / 2) * W)
            y0, y1 = int((cy[t] - h[t] / 2) * H), int((cy[t] + h[t] / 2) * H)
            masks[t, y0:y1, x0:x1] = True
        targets.append({"labels": torch.zeros(T, dtype=torch.long, device=device), "boxes": boxes.to(device),
                        "masks": masks.to(device), "valid": torch.ones(T, dtype=torch.long, device=device),
                        "size": torch.as_tensor([H, W], device=device), "orig_size": torch.as_tensor([H, W], device=device),
                        "caption": "a person riding a bike on the left"})
    return targets


def run(argv=None):
    """Parse `argv` (None = sys.argv), run the configured workload and return the result line as a dict on
    rank 0 (None on the other ranks; an int exit code for --stage).  bench.py --workload model3|model4 calls this
    in-process under its own clock sampler."""
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", action="store_true")
    ap.add_argument("--config", type=int, default=3, choices=(3, 4))
    ap.add_argument("--op", default="b200", choices=("b200", "reference"))
    ap.add_argument("--cpu-stub", action="store_true", help="CPU dry run of the harness (build container)")
    ap.add_argument("--frames", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--layers", type=int, default=None, help="enc_layers = dec_layers (config 3: 6 as BASELINE.json says; reference default 4)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--backbone", default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--module", default="reference", choices=("reference", "b200"),
                    help="which MSDeformAttn nn.Module class the reference transformer instantiates")
    ap.add_argument("--fused", action="store_true",
                    help="with --module b200: fused prologue (softmax + locations inside the gather kernel; encoder/FTF "
                         "layers skip materialising them) and fused value epilogue")
    ap.add_argument("--tf32", action="store_true",
                    help="torch.backends.cuda.matmul.allow_tf32 = True (NOT the reference's setting; shows what bounds the "
                         "model outside the hot path: its nn.Linear layers run as fp32 SIMT GEMMs by default)")
    ap.add_argument("--profile", default=None, help="write a torch.profiler table of one extra step (top CUDA kernels) to this file")
    ap.add_argument("--fused-topk", action="store_true",
                    help="decoder loop with this repo's one-launch sample selection (decoder_sample_topk) in place of the five "
                         "tensor ops of tce_deformable_transformer.py:752-758 (tools/decoder_topk_patch.py; no reference file is edited)")
    ap.add_argument("--compare-topk", action="store_true",
                    help="config 3: run the same clip through the reference's decoder loop and through the fused one and compare "
                         "the model outputs and the kept samples bit for bit (attention-weight ties are broken by a seeded bias "
                         "perturbation first: at initialisation ALL weights tie and torch.topk leaves that order unspecified)")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="config 3: capture the whole model forward in one CUDA graph and time replays (possible on this operator "
                         "only: the reference's forward calls cudaDeviceSynchronize, ms_deform_im2col_cuda.cuh:1260)")
    ap.add_argument("--compare", action="store_true",
                    help="after timing, run one more step on the OTHER operator in the same process (same weights, "
                         "same inputs) and report the largest relative difference of the model outputs")
    a = ap.parse_args(argv)
    if a.stage:
        return stage()

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu = a.cpu_stub
    device = torch.device("cpu" if cpu else f"cuda:{local}")
    if not cpu:
        torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("gloo" if cpu else "nccl")
    ext_name = install_environment(a.op, cpu)
    if a.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True

    production = ["--with_box_refine", "--binary", "--freeze_text_encoder", "--f_token", "8", "--qtrans",
                  "--device", str(device)]
    if a.config == 3:
        T, H, W = a.frames or 36, a.height or 360, a.width or 640
        layers = a.layers or 6
        backbone = a.backbone or "resnet50"
        n_clips = 1
    else:
        T, H, W = a.frames or 5, a.height or 384, a.width or 640       # 640 px long side, padded to /32 (level set B)
        layers = a.layers or 4
        backbone = a.backbone or "video_swin_t_p4w7"
        n_clips = 2
    flags = production + ["--backbone", backbone, "--enc_layers", str(layers), "--dec_layers", str(layers),
                          "--num_frames", str(T), "--batch_size", str(n_clips)]
    args = reference_args(flags)
    module_cls = swap_module_class() if (a.module == "b200" and not cpu) else None
    torch.manual_seed(1234)
    from models import build_model
    model, criterion, _ = build_model(args)
    model.to(device)
    n_fused = 0
    if module_cls is not None:
        for name, m in model.named_modules():
            if isinstance(m, module_cls):
                n_fused += 1
                if a.fused:
                    m.fused_prologue = True
                    m.fused_value_epilogue = True
                    m.check_shapes = "once"
                    # only the decoder's cross attention consumes the returned locations / weights
                    # (tce_deformable_transformer.py:752-758); encoder self-attention takes [0] (:543-544)
                    # and the FTF layer binds but never reads them (:451-452)
                    m.return_sampling = ".cross_attn" in name
        assert n_fused > 0, "the reference transformer did not pick up the swapped module class"
    topk_original = None
    if a.fused_topk and not cpu:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import decoder_topk_patch
        topk_original = decoder_topk_patch.install()
    calls = count_op_calls()
    n_params = sum(p.numel() for p in model.parameters())
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)

    g = torch.Generator().manual_seed(7 + rank)
    captions = ["a person riding a bike on the left"] * n_clips

    def sync():
        if not cpu:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    if a.config == 3:
        model.eval()
        clip = torch.randn(T, 3, H, W, generator=g).to(device)
        target = [{"size": torch.as_tensor([H, W], device=device)}]

        def step():
            with torch.no_grad():
                out = model([clip], captions, target)
            return out["pred_masks"]
    else:
        model.train()
        criterion.train()
        ddp = model
        if world > 1:
            ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=None if cpu else [local],
                                                            find_unused_parameters=False)       # main.py:67-72
        opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-5, weight_decay=5e-4)
        clips = [torch.randn(T, 3, H, W, generator=g).to(device) for _ in range(n_clips)]
        targets = synthetic_targets(n_clips, T, H, W, device, 11 + rank)
        from util.misc import nested_tensor_from_videos_list

        def step():
            samples = nested_tensor_from_videos_list(clips)
            out = ddp(samples, captions, targets)
            loss_dict = criterion(out, targets)
            wd = criterion.weight_dict
            loss = sum(loss_dict[k] * wd[k] for k in loss_dict if k in wd)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 0.1)                              # engine.py:77
            opt.step()
            return loss.detach()

    graph_note = None
    if a.cuda_graph and a.config == 3 and not cpu:
        eager_step = step
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = eager_step()

            def step():
                graph.replay()
                return static_out
            graph_note = "whole forward captured in one CUDA graph"
        except Exception as exc:      # what blocks the capture is the result
            step = eager_step
            torch.cuda.synchronize()
            graph_note = "capture failed: " + repr(exc)[:300]

    for _ in range(a.warmup):
        r = step()
    sync()
    calls["fwd"] = calls["bwd"] = 0
    calls["events"] = {"fwd": [], "bwd": []}
    b200_ext = None if (cpu or a.op != "b200") else importlib.import_module("tce-rvos_b200").load_extension()
    launches0 = b200_ext.launch_count() if b200_ext else 0
    t0 = time.perf_counter()
    if not cpu:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    for _ in range(a.steps):
        r = step()
    if not cpu:
        e1.record()
    sync()
    wall = time.perf_counter() - t0
    dev_s = wall if cpu else e0.elapsed_time(e1) / 1e3
    if world > 1:
        t = torch.tensor([dev_s], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s = float(t)
    frames = T * n_clips * world * a.steps
    b200_launches = (b200_ext.launch_count() - launches0) / a.steps if b200_ext else None
    if a.profile and not cpu and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        with open(a.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
    compare = None
    if a.compare and not cpu and a.config == 3:
        fn = importlib.import_module("models.ops.functions.ms_deform_attn_func")
        with torch.no_grad():
            mine = model([clip], captions, target)
        other = "reference" if a.op == "b200" else "b200"
        if other == "b200":
            other_ext = importlib.import_module("tce-rvos_b200").load_extension()
        else:
            spec = importlib.util.spec_from_file_location("MSDA_reference_cuda", os.path.join(ROOT, "baseline", "_ref", "MSDA_reference_cuda.so"))
            other_ext = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(other_ext)
        fn.MSDA = other_ext
        with torch.no_grad():
            theirs = model([clip], captions, target)
        compare = {"other_op": other}
        for k in ("pred_logits", "pred_boxes", "pred_masks"):
            x, y = mine[k].double(), theirs[k].double()
            compare[k] = float((x - y).abs().max() / y.abs().max())
    topk_compare = None
    if a.compare_topk and not cpu and a.config == 3:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import decoder_topk_patch
        if topk_original is not None:
            decoder_topk_patch.uninstall(topk_original)
        gen = torch.Generator(device="cpu").manual_seed(99)
        n_mod = 0
        for name, m in model.named_modules():
            if hasattr(m, "attention_weights") and hasattr(m, "sampling_offsets"):
                with torch.no_grad():
                    m.attention_weights.bias.add_(torch.randn(m.attention_weights.bias.shape, generator=gen).to(device))
                n_mod += 1
        kept = {}
        decoder = model.transformer.decoder
        hook = decoder.register_forward_hook(lambda mod, inp, out: kept.__setitem__("s", out[2].detach().clone()))
        with torch.no_grad():
            ref_out = model([clip], captions, target)
        ref_samples = kept["s"]
        original = decoder_topk_patch.install()
        with torch.no_grad():
            new_out = model([clip], captions, target)
        new_samples = kept["s"]
        decoder_topk_patch.uninstall(original)
        hook.remove()
        topk_compare = {"modules_with_perturbed_weight_bias": n_mod,
                        "kept_samples_shape": list(ref_samples.shape),
                        "kept_samples_bit_identical": bool(torch.equal(ref_samples, new_samples)),
                        "kept_samples_max_abs_diff": float((ref_samples - new_samples).abs().max())}
        for k in ("pred_logits", "pred_boxes", "pred_masks"):
            topk_compare[k + "_bit_identical"] = bool(torch.equal(ref_out[k], new_out[k]))
    line = {
        "what": f"BASELINE.json configs[{a.config - 1}]: unmodified reference model on the {a.op} operator",
        "op_extension": ext_name, "module_class": a.module + ("+fused" if a.fused else ""), "config": a.config, "backbone": backbone, "enc_dec_layers": layers,
        "frames_per_clip": T, "clips_per_gpu": n_clips, "input": [H, W], "mode": "inference" if a.config == 3 else "train step (fwd+loss+bwd+clip+AdamW)",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dev_s / a.steps * 1e3,
        "frames_per_s": frames / dev_s, "op_calls_per_step": {k: calls[k] / a.steps for k in ("fwd", "bwd")},
        "op_ms_per_step": {k: sum(x.elapsed_time(y) for x, y in v) / a.steps for k, v in calls["events"].items()},
        "b200_kernel_launches_per_step": b200_launches, "params_M": n_params / 1e6, "trainable_params_M": n_train / 1e6,
        "result_checksum": float(r.float().abs().mean()) if hasattr(r, "float") else None,
        "weights": "random init (seed 1234), synthetic clip", "dtype": "fp32" + (" (matmul TF32 allowed)" if a.tf32 else ""), "compare_max_rel": compare,
        "decoder_topk": "fused (decoder_sample_topk)" if a.fused_topk else "reference loop", "decoder_topk_compare": topk_compare,
        "cuda_graph": graph_note,
    }
    if rank == 0 and a.out:
        with open(a.out, "a") as f:
            f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


def main():
    r = run()
    if isinstance(r, int):
        return r
    if r is not None:
        print(json.dumps(r))
    return 0


if __name__ == "__main__":
    sys.exit(main())
