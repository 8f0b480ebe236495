"""Camera-frame front end on the GPU (d3_frame_to_input / d3_depth_to_input, d3roma_b200.inference.D3RoMa) against the
cv2-generated golden vectors and the oracle's restatement of inference.py:97-250.  Integer / byte work: bit-exact."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden" / "frontend.npz"
CASES = ["down", "up", "same", "odd", "gray"]


@pytest.mark.parametrize("name", CASES)
def test_frame_and_depth_kernels_bit_exact_vs_cv2_golden(name):
    import d3roma_b200 as d
    from d3roma_b200 import inference as I
    from oracle import frontend_ref as F
    g = np.load(GOLD)
    h, w = (int(v) for v in g[f"{name}/size"])
    padder = d.InputPadder((1, 1, h, w), divis_by=32)
    normer = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
    img = torch.from_numpy(g[f"{name}/img"])
    x = I.frames_to_input(torch.stack([img, img.flip(0)]), (h, w), padder, "cuda")      # batch of two frames
    want = F.pad_replicate(torch.from_numpy(g[f"{name}/normalized"]), 32)
    assert x.shape == (2, 3) + tuple(want.shape[-2:])
    assert torch.equal(x[0].cpu(), want)
    want1 = F.pad_replicate(F.preprocess_image(g[f"{name}/img"][::-1].copy(), (h, w)), 32) if True else None
    assert torch.equal(x[1].cpu(), want1)
    raw, mask, norm = I.depth_to_input(g[f"{name}/depth"], (h, w), padder, 24.547, normer, "cuda")
    assert torch.equal(raw[0].cpu(), torch.from_numpy(g[f"{name}/raw_disp"]))
    assert torch.equal(mask[0].cpu(), torch.from_numpy((g[f"{name}/raw_disp"] > 0).astype(np.float32)))
    assert torch.equal(norm[0].cpu(), F.pad_replicate(torch.from_numpy(g[f"{name}/norm_disp"]), 32))


def test_full_size_frames_match_cv2_directly():
    """848x480 RealSense-sized frames -> 640x360: bit-exact against cv2 run here (skipped without OpenCV)."""
    cv2 = pytest.importorskip("cv2")
    import d3roma_b200 as d
    from d3roma_b200 import inference as I
    from oracle import frontend_ref as F
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (480, 848, 3), dtype=np.uint8)
    depth = (rng.random((480, 848)) * 2).astype(np.float32)
    padder = d.InputPadder((1, 1, 360, 640), divis_by=32)
    normer = d.Normalizer(ssi=False, mode="average", num_chs=1, ch_bounds=[128.], ch_gammas=[1.], t=0.5, s=2.0)
    x = I.frames_to_input(img, (360, 640), padder, "cuda")
    assert torch.equal(x[0].cpu(), F.pad_replicate(F.preprocess_image(img, (360, 640)), 32))
    raw, mask, norm = I.depth_to_input(depth, (360, 640), padder, 24.547, normer, "cuda")
    r_raw, r_mask, r_norm = F.preprocess_depth(depth, (360, 640), 24.547)
    assert torch.equal(raw[0].cpu(), r_raw) and torch.equal(mask[0].cpu(), r_mask)
    assert torch.equal(norm[0].cpu(), F.pad_replicate(r_norm, 32))


def test_d3roma_infer_matches_reference_flow():
    """D3RoMa.infer (inference.py:137-250): numpy frames in, depth in metres out; equals the reference flow assembled
    from the oracle's preprocessing + the same pipeline + torch post-processing."""
    pytest.importorskip("cv2")
    import d3roma_b200 as d
    from d3roma_b200 import evaluate
    from d3roma_b200.inference import D3RoMa
    from d3roma_b200.weights import random_init_state_dict
    from oracle import frontend_ref as F
    h, w = 60, 90
    net = d.UNet2DModel(sample_size=(64, 96), in_channels=8, out_channels=1, compute_dtype="fp16").to("cuda")
    net.load_state_dict(random_init_state_dict(8, 1, seed=0))
    sched = d.MyDDIMScheduler.from_config(dict(beta_schedule="squaredcos_cap_v2", num_train_timesteps=128,
                                               prediction_type="sample", clip_sample=True, set_alpha_to_one=False,
                                               steps_offset=1))
    pipe = d.GuidedDiffusionPipeline(net, sched, d.FlowGuidance(flow_guidance_weight=0.0))
    cfg = evaluate.default_config(num_inference_timesteps=2, num_inference_rounds=2, camera_resolution=f"{w}x{h}")
    cam = d.SIM_CAMERA
    rng = np.random.default_rng(3)
    left = rng.integers(0, 256, (120, 180), dtype=np.uint8)            # gray IR frames, 2x the working resolution
    right = rng.integers(0, 256, (120, 180), dtype=np.uint8)
    rgb = rng.integers(0, 256, (120, 180, 3), dtype=np.uint8)
    depth = (rng.random((120, 180)) * 1.5 + 0.3).astype(np.float32)
    depth[rng.random((120, 180)) < 0.2] = 0
    eng = D3RoMa(pipe, cam, cfg)
    torch.manual_seed(11)
    got = eng.infer(left, right, depth, rgb)
    assert got.shape == (h, w) and got.dtype == np.float32 and np.isfinite(got).all()
    assert set(eng.last_metrics) == {"disp", "depth"}
    # reference flow
    R = 2
    pad = lambda t: F.pad_replicate(t, 32)[None].repeat(R, 1, 1, 1).cuda()   # noqa: E731
    n_l, n_r, n_rgb = (pad(F.preprocess_image(a, (h, w))) for a in (left, right, rgb))
    raw, mask, norm = F.preprocess_depth(depth, (h, w), cam.fxb)
    torch.manual_seed(11)
    out = pipe(n_rgb, n_l, n_r, pad(norm), None, None, num_inference_steps=2, num_intermediate_images=1,
               depth_channels=1, cond_channels="left+right+raw")
    pred = out.images[:, :, 2:62, 3:93].cpu()
    disp = ((pred / 2.0 + 0.5) * 128.0).mean(dim=0)[0]
    want = torch.where(disp > 0, cam.fxb / disp, torch.zeros_like(disp)).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)


def test_d3roma_from_config_loads_the_saved_pipeline(tmp_path):
    """D3RoMa.from_config = the reference's constructor path (inference.py:40-95): pipeline directory named by
    config.resume_pretrained, class picked by config.ldm, my_ddim sampler override, guidance mode from the config."""
    import d3roma_b200 as d
    from d3roma_b200 import evaluate
    from d3roma_b200.inference import D3RoMa
    from d3roma_b200.weights import random_init_state_dict
    net = d.UNet2DModel(sample_size=(64, 96), in_channels=8, out_channels=1, compute_dtype="fp16").to("cuda")
    net.load_state_dict(random_init_state_dict(8, 1, seed=0))
    ddpm = d.MyDDPMScheduler.from_config(dict(beta_schedule="squaredcos_cap_v2", num_train_timesteps=128,
                                              prediction_type="sample"))
    d.GuidedDiffusionPipeline(net, ddpm, d.FlowGuidance(flow_guidance_weight=0.0)).save_pretrained(tmp_path / "ckpt")
    cfg = evaluate.default_config(num_inference_timesteps=2, camera_resolution="90x60", ldm=False,
                                  resume_pretrained=str(tmp_path / "ckpt"), flow_guidance_mode="gradient",
                                  sampler="my_ddim", beta_schedule="squaredcos_cap_v2", beta_start=1e-4, beta_end=2e-2,
                                  clip_sample=True, num_train_timesteps=128, prediction_type="sample")
    eng = D3RoMa.from_config(cfg, d.SIM_CAMERA)
    assert type(eng.pipeline.scheduler).__name__ == "MyDDIMScheduler" and eng.pipeline.scheduler.config.steps_offset == 1
    assert eng.pipeline.guidance.flow_guidance_mode == "gradient"
    rng = np.random.default_rng(4)
    rgb = rng.integers(0, 256, (60, 90, 3), dtype=np.uint8)
    gray = rng.integers(0, 256, (60, 90), dtype=np.uint8)
    depth = (rng.random((60, 90)) + 0.3).astype(np.float32)
    got = eng.infer(gray, gray, depth, rgb)
    assert got.shape == (60, 90) and np.isfinite(got).all()
    cfg.resume_pretrained = str(tmp_path / "missing")
    with pytest.raises(ValueError):
        D3RoMa.from_config(cfg, d.SIM_CAMERA)


def test_depth_to_rgb_frame_reprojection_bit_exact_vs_reference_and_oracle():
    """DepthCamera.transform_depth_to_rgb_frame / transform_cropped_depth_to_rgb_frame (utils_d3roma/camera.py:610-701)
    on the device: bit-identical to the reference camera object's outputs (tests/golden/reproject.npz, including numpy's
    last-writer-wins collision order and the 4-corner aggressive fill) and to the oracle on a larger random map."""
    from pathlib import Path
    import numpy as np
    from d3roma_b200.camera import DepthCamera
    from oracle import reproject_ref as R
    z = np.load(Path(__file__).resolve().parent / "golden" / "reproject.npz")
    H, W = (int(v) for v in z["hw"])
    cam = DepthCamera(f"{W}x{H}", 1.0, intrinsics={"color": z["K_color"], "depth": z["K_depth"]},
                      extrinsics={"color_to_depth": z["color_to_depth"]}, T_fc=z["T_fc"],
                      config={"auto_scale": True, "aggressive_fill": True})
    assert np.allclose(cam.T_cl, z["T_cl"])
    depth = z["depth_mm"].astype(np.float32) / np.float32(1000.0)
    mm = lambda o: np.round(o.astype(np.float64) * 1000.0).astype(np.uint16)   # noqa: E731
    out = cam.transform_depth_to_rgb_frame(depth)
    assert isinstance(out, np.ndarray) and np.array_equal(mm(out), z["out_aggressive_mm"])
    cam.config["aggressive_fill"] = False
    out_t = cam.transform_depth_to_rgb_frame(torch.from_numpy(depth).cuda())        # device in, device out
    assert out_t.is_cuda and np.array_equal(mm(out_t.cpu().numpy()), z["out_round_mm"])
    oy, ox, ch, cw = (int(v) for v in z["crop"])
    out_c = cam.transform_cropped_depth_to_rgb_frame(depth[oy:oy + ch, ox:ox + cw], z["cropped_K"])
    assert np.array_equal(mm(out_c), z["out_cropped_mm"])
    # random depth (every value distinct: exposes any difference in collision order) against the oracle, both fills
    rng = np.random.default_rng(3)
    d2 = (rng.random((H, W)) * 1.5 + 0.25).astype(np.float32)
    d2[rng.random((H, W)) < 0.2] = 0.0
    for aggressive in (True, False):
        cam.config["aggressive_fill"] = aggressive
        want = R.reproject(d2, cam.K_depth.inv, cam.T_cl, cam.K_color.arr, (H, W), aggressive)
        assert np.array_equal(cam.transform_depth_to_rgb_frame(d2), want)
    # identity extrinsics: returned unchanged, like the reference (:655-657)
    cam_id = DepthCamera(f"{W}x{H}", 1.0, intrinsics={"color": z["K_color"], "depth": z["K_depth"]},
                         extrinsics={"color_to_depth": np.eye(4)})
    assert cam_id.transform_depth_to_rgb_frame(depth) is depth
