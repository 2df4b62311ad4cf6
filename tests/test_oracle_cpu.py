"""CPU tests of the ORACLE (test infrastructure) against the committed golden fixtures.

The fixtures under tests/golden/ were written by oracle/make_golden.py in the build container: it imports the
reference's own unmodified modules from /root/reference (through oracle/shims), runs StableMTLPipeline.single_infer
for every task on seeded inputs/weights and stores (a) the reference's outputs, (b) the oracle's fp32 outputs.
Here -- with no access to /root/reference -- the oracle is re-run on the same seeds and must reproduce both."""
import os

import pytest
import torch

from oracle import stablemtl_oracle as O
from stablemtl_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(fx, ucfg, vcfg):
    s = fx["seeds"]
    child = synth.make_unet_state_dict(ucfg, seed=s["child"])
    vae = synth.make_vae_state_dict(vcfg, seed=s["vae"])
    text = synth.make_text_embeddings(ucfg.cross_attention_dim, seed=s["text"])
    main = None
    if fx["multi"]:
        main = dict(synth.make_unet_state_dict(ucfg, seed=s["main"]))
        main.update(synth.make_task_modules_state_dict(ucfg, seed=s["task"]))
    rgb, nxt = synth.make_images(fx["batch"], fx["h"], fx["w"], seed=fx["image_seed"])
    return O.Oracle(ucfg, vcfg, child, vae, text, main), rgb, nxt


@pytest.mark.parametrize("name", ["tiny_single_64x96", "tiny_single_40x72"])
def test_oracle_reproduces_golden_single_stream(name):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    fx = torch.load(os.path.join(GOLDEN, name + ".pt"))
    assert fx["pin_float64_rel_l2"] <= 1e-7           # the pin recorded when the fixture was made
    orc, rgb, nxt = build(fx, synth.TINY_UNET, synth.TINY_VAE)
    maps, clipped, latents = orc.predict_all(rgb, nxt, return_latents=True)
    for t in synth.TASKS:
        # same algorithm, same seeds, fp32: only thread-count-dependent summation order may differ
        assert rel_l2(clipped[t], fx["oracle_fp32_clipped"][t]) <= 1e-4, t
        assert rel_l2(latents[t], fx["oracle_fp32_latents"][t]) <= 1e-4, t
        # the reference's own output (its self-attention inputs are cast to fp16, attention.py:392-394)
        assert rel_l2(clipped[t], fx["reference_clipped"][t]) <= 3e-3, t
    agree = (maps["semantic"] == fx["oracle_fp32_semantic"]).float().mean().item()
    assert agree >= 0.9995


def test_oracle_reproduces_golden_sd2_multi_stream():
    """SD-2-sized UNet/VAE, multi-stream (7 child passes + task attention), 32x48 image: the reference's outputs."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    fx = torch.load(os.path.join(GOLDEN, "sd2_multi_32x48.pt"))
    assert fx["pin_float64_rel_l2"] <= 1e-7
    orc, rgb, nxt = build(fx, synth.SD2_UNET, synth.SD2_VAE)
    # two tasks keep the CPU suite short: a 1-channel mean task and a flow task (second frame encoded)
    rgb_n, nxt_n = rgb / 255.0 * 2.0 - 1.0, nxt / 255.0 * 2.0 - 1.0
    cache = {}
    for t in ("depth", "optical_flow"):
        out = orc.single_infer(rgb_n, nxt_n, t, cache)
        assert rel_l2(out, fx["oracle_fp32_clipped"][t]) <= 1e-4, t
        assert rel_l2(out, fx["reference_clipped"][t]) <= 3e-3, t


def test_task_attention_is_not_vacuous():
    """to_out_task is zero-initialised in the reference (util/model.py:145-146); the synthetic checkpoint
    re-randomises it so the multi-stream parity tests exercise the task branch.  Zeroing it must change the output."""
    ucfg, vcfg = synth.TINY_UNET, synth.TINY_VAE
    child = synth.make_unet_state_dict(ucfg, seed=0)
    vae = synth.make_vae_state_dict(vcfg, seed=2)
    text = synth.make_text_embeddings(ucfg.cross_attention_dim)
    main = dict(synth.make_unet_state_dict(ucfg, seed=10))
    main.update(synth.make_task_modules_state_dict(ucfg, seed=11))
    rgb, nxt = synth.make_images(1, 32, 48, seed=3)
    rn, nn = rgb / 255.0 * 2.0 - 1.0, nxt / 255.0 * 2.0 - 1.0
    a = O.Oracle(ucfg, vcfg, child, vae, text, main).single_infer(rn, nn, "normal")
    zeroed = {k: (torch.zeros_like(v) if ".to_out_task." in k else v) for k, v in main.items()}
    b = O.Oracle(ucfg, vcfg, child, vae, text, zeroed).single_infer(rn, nn, "normal")
    assert rel_l2(a, b) > 1e-3
    # and the main task's own child stream is excluded (stablemtl_pipeline.py:483-484): the output for "normal"
    # cannot depend on the child's "normal" text embedding
    text2 = dict(text)
    text2["normal"] = text["normal"]            # unchanged for the main pass
    orc = O.Oracle(ucfg, vcfg, child, vae, text2, main)
    cache = {}
    feats = orc.child_features(rn, nn, cache)
    feats["normal"] = [torch.randn_like(f) for f in feats["normal"]]
    c = orc.single_infer(rn, nn, "normal", cache)
    assert rel_l2(c, a) <= 1e-6


def test_postprocess_conventions():
    """stablemtl_pipeline.py:297-366."""
    x = torch.tensor([[[[0.5]], [[-0.5]], [[1.0]]]])
    assert torch.allclose(O.postprocess(x, "albedo"), (x + 1) / 2)
    n = O.postprocess(x, "normal")
    assert torch.allclose(n.norm(dim=1), torch.ones(1, 1, 1))
    z = torch.zeros(1, 3, 1, 1)
    assert torch.equal(O.postprocess(z, "normal"), z)                  # zero vector stays zero (norm -> 1)
    pal = torch.tensor(O.PALETTE, dtype=torch.float32) / 255.0 * 2.0 - 1.0
    sem = O.postprocess(pal.t().reshape(1, 3, 1, 8), "semantic")
    assert sem.flatten().tolist() == list(range(8))
    assert O.select_channels(torch.ones(2, 3, 4, 4), "depth").shape == (2, 1, 4, 4)
    assert O.select_channels(torch.ones(2, 3, 4, 4), "optical_flow").shape == (2, 2, 4, 4)


def test_explicit_size_nearest_upsample_rule():
    """Appendix A: legacy nearest with explicit size, src = floor(dst * in / out); 8 -> 15 as in 8x10 -> 15x20."""
    x = torch.arange(8.0).reshape(1, 1, 8, 1)
    y = torch.nn.functional.interpolate(x, size=(15, 1), mode="nearest").flatten().tolist()
    assert y == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7]
