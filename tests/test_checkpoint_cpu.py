"""Checkpoint ingest from the reference's on-disk layout (synthetic files written in that layout)."""
import os

import pytest
import torch

from stablemtl_b200 import checkpoint as ck
from stablemtl_b200 import synth


def _write_layout(tmp, multi):
    ucfg, vcfg = synth.TINY_UNET, synth.TINY_VAE
    sd2 = tmp / "base" / "stable-diffusion-2"
    (sd2 / "unet").mkdir(parents=True)
    (sd2 / "vae").mkdir(parents=True)
    unet = synth.make_unet_state_dict(ucfg, seed=0)
    sd2_unet = dict(unet)
    sd2_unet["conv_in.weight"] = unet["conv_in.weight"][:, :4].clone()            # SD-2 ships a 4-channel conv_in
    torch.save(sd2_unet, sd2 / "unet" / "diffusion_pytorch_model.bin")
    vae = synth.make_vae_state_dict(vcfg, seed=2)
    legacy = {}
    for k, v in vae.items():                                                      # write the legacy attention names
        for new, old in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
            if f".attentions.0.{new}." in k:
                k = k.replace(f".attentions.0.{new}.", f".attentions.0.{old}.")
                if k.endswith("weight"):
                    v = v[:, :, None, None]
        legacy[k] = v
    torch.save(legacy, sd2 / "vae" / "diffusion_pytorch_model.bin")
    run = tmp / "run"
    (run / "checkpoint" / "latest" / "unet").mkdir(parents=True)
    main = dict(synth.make_unet_state_dict(ucfg, seed=10))
    if multi:
        main.update(synth.make_task_modules_state_dict(ucfg, seed=11))
    torch.save(main, run / "checkpoint" / "latest" / "unet" / "diffusion_pytorch_model.bin")
    child_path = tmp / "single_stream_unet.pth"
    torch.save(unet, child_path)
    return sd2_unet, unet, vae, main, str(tmp / "base"), str(run), str(child_path)


def test_single_stream_ingest(tmp_path):
    sd2_unet, unet, vae, main, base, run, child_path = _write_layout(tmp_path, multi=False)
    child, m, v = ck.load_reference_checkpoints(base, run_dir=run)
    assert m is None and all(torch.equal(child[k], main[k]) for k in main)
    assert set(v) == set(vae) and all(torch.equal(v[k], vae[k]) for k in vae)     # legacy names / conv-shaped weights undone
    child, m, _ = ck.load_reference_checkpoints(base)                              # plain SD-2: conv_in widened 4 -> 12
    w4 = sd2_unet["conv_in.weight"]
    assert child["conv_in.weight"].shape == (w4.shape[0], 12, 3, 3)
    assert torch.allclose(child["conv_in.weight"], w4.repeat(1, 3, 1, 1) / 3)      # src/util/model.py:14-15
    x = torch.randn(1, 4, 5, 5)                                                    # duplicating the input keeps the conv's output
    assert torch.allclose(torch.nn.functional.conv2d(x.repeat(1, 3, 1, 1), child["conv_in.weight"]),
                          torch.nn.functional.conv2d(x, w4), atol=1e-5)


def test_multi_stream_ingest(tmp_path):
    _, unet, vae, main, base, run, child_path = _write_layout(tmp_path, multi=True)
    child, m, v = ck.load_reference_checkpoints(base, run_dir=run, single_stream_path=child_path)
    assert all(torch.equal(child[k], unet[k]) for k in unet)
    assert any(".task_to_q." in k for k in m) and all(torch.equal(m[k], main[k]) for k in main)
    with pytest.raises(ValueError, match="task modules"):
        ck.load_reference_checkpoints(base, single_stream_path=child_path)         # no trained run: nothing to attend with
    with pytest.raises(FileNotFoundError):
        ck.load_reference_checkpoints(os.path.join(base, "missing"))


def test_rejects_pickled_objects(tmp_path):
    p = tmp_path / "x.bin"
    torch.save({"a": [1, 2, 3]}, p)
    with pytest.raises(Exception):
        ck.load_tensor_file(str(p))
