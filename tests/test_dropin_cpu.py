"""Host logic of the drop-in facade (stablemtl_b200/dropin.py) against a stub engine: the attribute / call sequence of
the reference trainer (src/trainer/stablemtl_trainer.py:405-437, 262-305) must go through without a GPU-side surprise.
The numerics behind it are checked on the GPU (tests/test_pipeline_gpu.py::test_dropin_*)."""
import pytest
import torch

from stablemtl_b200 import synth
from stablemtl_b200.dropin import StableMTLPipeline, _ModuleProxy


class StubWeights:
    def __init__(self, text, tasks):
        self.ntok = [text[t].shape[0] for t in tasks]
        self.text = torch.zeros(len(tasks), 4, text[tasks[0]].shape[1])
        for i, t in enumerate(tasks):
            self.text[i, : self.ntok[i]] = text[t]


class StubEngine:
    """records the stage calls; shapes follow the real engine"""
    device = torch.device("cpu")

    def __init__(self, multi=True):
        self.multi, self.tasks, self.ucfg = multi, list(synth.TASKS), synth.TINY_UNET
        self.child_w = StubWeights(synth.make_text_embeddings(self.ucfg.cross_attention_dim), self.tasks)
        self.calls = []
        self.last = {}

    def encode_rgb(self, x):
        self.calls.append(("encode_rgb", tuple(x.shape)))
        return x[:, :1].repeat(1, 4, 1, 1)[:, :, ::8, ::8] * (1.0 + x.mean())

    def decode_latents(self, lat):
        self.calls.append(("decode", tuple(lat.shape)))
        B, _, h, w = lat.shape
        return torch.arange(3.0).view(1, 3, 1, 1).expand(B, 3, 8 * h, 8 * w).clone()

    from stablemtl_b200.pipeline import StableMTLEngine as _E
    task_of_text = _E.task_of_text

    def unet_forward(self, which, sample, task, task_feats=None):
        self.calls.append(("unet", which, task, None if task_feats is None else sorted(task_feats[0])))
        B, _, _, h, w = sample.shape
        n = len(self.ucfg.transformer_dims())
        taps = [torch.zeros(B, h * w, 8) for _ in range(n)] if which == "child" else None
        return torch.zeros(B, 4, 1, h, w), taps

    def predict(self, rgb, nxt=None):
        self.calls.append(("predict", float(rgb.sum()), None if nxt is None else float(nxt.sum())))
        B, _, H, W = rgb.shape
        self.last = {t: torch.full((B, 3, H, W), float(len(self.calls))) for t in self.tasks}


def accelerator_prepare(*objs):
    """accelerate.Accelerator.prepare leaves anything that is not a Module / Optimizer / DataLoader / scheduler as it is"""
    for o in objs:
        assert not isinstance(o, torch.nn.Module)
    return objs if len(objs) > 1 else objs[0]


def test_trainer_eval_preamble_runs_against_the_facade():
    pipe = StableMTLPipeline(StubEngine(multi=True))
    dev = torch.device("cpu")
    # stablemtl_trainer.py:418-434
    pipe.vae.to(dev)
    pipe.text_encoder.to(dev)
    comps = [pipe.unet]
    assert pipe.unet_child is not None
    comps.insert(1, pipe.unet_child)
    pipe.unet, pipe.unet_child = accelerator_prepare(*comps)
    pipe.unet.eval()
    assert list(pipe.unet.parameters()) == [] and pipe.scheduler.config.prediction_type == "sample"
    assert pipe.rgb_latent_scale_factor == pipe.latent_scale_factor == 0.18215
    with pytest.raises(RuntimeError):
        pipe.unet.train()
    single = StableMTLPipeline(StubEngine(multi=False))
    assert single.unet_child is None and accelerator_prepare(single.unet) is single.unet
    assert single.create_task_feats(None, None, None, "depth", synth.TASKS, None, True) == (None, None)


def test_stage_methods_follow_the_reference_shapes_and_branches():
    eng = StubEngine(multi=True)
    pipe = StableMTLPipeline(eng)
    B, H, W = 2, 64, 96
    rgb, nxt = torch.rand(B, 3, H, W) * 2 - 1, torch.rand(B, 3, H, W) * 2 - 1
    lat = pipe.encode_rgb_latent("depth", rgb, nxt)                        # duplicate: one encode
    assert lat.shape == (B, 8, 1, H // 8, W // 8) and torch.equal(lat[:, :4], lat[:, 4:])
    assert [c[0] for c in eng.calls] == ["encode_rgb"]
    lat = pipe.encode_rgb_latent("optical_flow", rgb, nxt)                 # flow task with a next frame: two encodes
    assert not torch.equal(lat[:, :4], lat[:, 4:]) and len(eng.calls) == 3
    zero = StableMTLPipeline(StubEngine(), encode_rgb_model="zero").encode_rgb_latent("normal", rgb, None)
    assert float(zero[:, 4:].abs().max()) == 0.0
    with pytest.raises(AssertionError):
        pipe.encode_rgb_latent("bogus", rgb, nxt)
    with pytest.raises(ValueError):
        StableMTLPipeline(StubEngine(), encode_rgb_model="avg")
    text = pipe.create_text_condition(["scene_flow"], B)
    assert text.shape == (B, 4, eng.ucfg.cross_attention_dim)
    ts = torch.ones(B, dtype=torch.long) * 999
    eng.calls.clear()
    outs, feats = pipe.create_task_feats(rgb, nxt, ts, output_type="depth", task_output_types=synth.TASKS,
                                         rand_num_generator=None, drop_ratio=0.0, exclude_mainstream_output_type=True)
    assert len(outs) == 6 and outs[0].shape == (B, 4, 1, H // 8, W // 8) and len(feats) == 16
    assert sorted(feats[0]) == sorted(t for t in synth.TASKS if t != "depth")
    assert [c[2] for c in eng.calls if c[0] == "unet"] == [i for i, t in enumerate(synth.TASKS) if t != "depth"]
    cat = torch.cat([pipe.encode_rgb_latent("depth", rgb, nxt), torch.zeros(B, 4, 1, H // 8, W // 8)], 1)
    out, ret = pipe.unet(cat, ts, pipe.create_text_condition(["depth"], B), task_feats=feats, output_type="depth")
    assert out.sample.shape == (B, 4, 1, H // 8, W // 8) and len(ret) == 16 and ret[0] is None
    assert eng.calls[-1][:3] == ("unet", "main", synth.TASKS.index("depth"))
    with pytest.raises(ValueError):                                        # the time embedding is folded at t = 999
        pipe.unet(cat, torch.ones(B, dtype=torch.long) * 500, text)
    with pytest.raises(ValueError):                                        # not one of the constant prompts
        pipe.unet(cat, ts, torch.randn(B, 3, eng.ucfg.cross_attention_dim), task_feats=feats)
    with pytest.raises(ValueError):                                        # prompt / output_type mismatch
        pipe.unet(cat, ts, pipe.create_text_condition(["normal"], B), task_feats=feats, output_type="depth")
    d = pipe.decode_output(torch.zeros(B, 4, 8, 12), "depth")
    assert d.shape == (B, 1, 64, 96) and float(d.mean()) == 1.0
    assert pipe.decode_output(torch.zeros(B, 4, 8, 12), "optical_flow").shape == (B, 2, 64, 96)
    assert pipe.decode_output(torch.zeros(B, 4, 8, 12), "semantic").shape == (B, 3, 64, 96)
    with pytest.raises(ValueError):
        pipe.decode_output(torch.zeros(B, 4, 8, 12), "bogus")


def test_result_cache_is_keyed_on_the_content_of_both_frames():
    eng = StubEngine(multi=True)
    pipe = StableMTLPipeline(eng)
    rgb, nxt = torch.rand(1, 3, 64, 96) * 2 - 1, torch.rand(1, 3, 64, 96) * 2 - 1
    kw = dict(num_inference_steps=1, generator=None, show_pbar=False, exclude_mainstream_output_type=True,
              task_output_types=synth.TASKS)
    a = pipe.single_infer(rgb, output_type="depth", rgb_next_norm=nxt, **kw)
    b = pipe.single_infer(rgb.clone(), output_type="normal", rgb_next_norm=nxt.clone(), **kw)   # same content: served
    assert sum(c[0] == "predict" for c in eng.calls) == 1 and torch.equal(a, b)
    pipe.single_infer(rgb, output_type="depth", rgb_next_norm=nxt + 0.01, **kw)                  # another next frame
    assert sum(c[0] == "predict" for c in eng.calls) == 2
    pipe.single_infer(rgb, output_type="depth", rgb_next_norm=None, **kw)
    assert sum(c[0] == "predict" for c in eng.calls) == 3
    with pytest.raises(ValueError):
        pipe.single_infer(rgb, output_type="depth", rgb_next_norm=nxt, num_inference_steps=1, generator=None,
                          show_pbar=False, exclude_mainstream_output_type=True, task_output_types=["depth", "normal"])
    with pytest.raises(ValueError):
        pipe.single_infer(rgb, output_type="bogus", rgb_next_norm=nxt, **kw)


def test_proxies_are_not_modules_and_survive_device_moves():
    p = _ModuleProxy(StubEngine())
    assert p.to("cpu") is p and p.eval() is p and p.requires_grad_(False) is p and p.device == torch.device("cpu")
    with pytest.raises(RuntimeError):
        p.state_dict()
