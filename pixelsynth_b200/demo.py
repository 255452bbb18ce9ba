"""Drop-in for the reference's demo.py command line (demo.py:1-262, scripts/demo_image.sh, scripts/demo_scene.sh):

    python -m pixelsynth_b200.demo --vqvae --use_fixed_testset --model_setting gen_img \
        --old_model modelcheckpoints/realestate/pixelsynth.pth --gpu 0 --demo_img_name 1011.png \
        --result_folder demo/1011 --temperature=.7 --num_samples 1 --direction L --rotation .6

Same flags (options/test_options.py:12-197, the ones the demo path reads), same option merge (utils/opts_helper.py:3-55),
same input transform and cameras (demo.py:27-98), same output files (demo.py:100-178, 246-262).  The model behind it is
pixelsynth_b200.models.ZbufferModelPts on the sm_100a kernels: there is no CPU path, so running it needs a B200.

Checkpoint layout (demo.py:198-221): torch.load(old_model) = {"state_dict": BaseModel.state_dict() -- keys
`model.module.<submodule>.<param>` (+ `netD.*`), "opts": the training Namespace}; buffers named `xyzs` / `ones` are
skipped (the pixel grid is rebuilt for the requested size); `--load_vqvae` / `--load_autoregressive` overlay
separately trained VQ-VAE-2 (`module.` prefixed) and PixelCNN (`model_state_dict`) weights.  Without --old_model the
networks are random-initialised from --seed (pixelsynth_b200/synthetic.py) and the options are the shipped
RealEstate10K configuration (scripts/train_dpr_realestate.sh:5-18): no trained checkpoint is reachable offline.
"""
import argparse
import os
import types

import numpy as np
import torch

DEFAULT_OPTS = dict(  # scripts/train_dpr_realestate.sh:5-18 + options/train_options.py defaults + options.py:68-70
    W=256, splatter="xyblending", learn_default_feature=True, radius=4.0, pp_pixel=128, rad_pow=2, tau=1.0,
    accumulation="alphacomposite", min_z=1.0, max_z=100.0, use_rgb_features=True, use_gt_depth=False,
    use_inverse_depth=False, depth_predictor_type="unet", Unet_num_filters=32, refine_model_type="resnet_256W8UpDown3",
    ngf=64, norm_G="sync:spectral_batch", predict_residual=True, normalize_image=True, dataset="realestate", seed=0)


def build_parser():
    """The subset of options/test_options.py:12-197 the demo reads; every flag keeps its name, type and default.
    (`--gpu 0,1` in the shipped scripts is argparse's unambiguous abbreviation of --gpu_ids; it stays one here.)"""
    p = argparse.ArgumentParser(prog="pixelsynth_b200.demo")
    a = p.add_argument
    a("--old_model", type=str, default="")
    a("--result_folder", type=str, default="")
    a("--model_setting", type=str, default="train",
      choices=("train", "gen_paired_img", "gen_img", "gen_scene", "get_gen_order", "gen_two_imgs"))
    a("--dataset_folder", type=str, default="")
    a("--demo_img_name", type=str, default="")
    a("--num_samples", type=int, default=1)
    a("--temperature", type=float, default=1.0)
    a("--temp_eps", type=float, default=0.05)
    a("--rotation", type=float, default=0.3)
    a("--decoder_truncation_threshold", type=float, default=2)
    a("--homography", action="store_true", default=False)
    a("--load_autoregressive", action="store_true", default=False)
    a("--no_outpainting", action="store_true", default=False)
    a("--render_ids", type=int, nargs="+", default=[1])
    a("--directions", type=str, nargs="+", default=[])
    a("--direction", type=str, default="")
    a("--background_smoothing_kernel_size", type=int, default=13)
    a("--normalize_before_residual", action="store_true", default=False)
    a("--sequential_outpainting", action="store_true", default=False)
    a("--pretrain", action="store_true", default=False)
    a("--val_rotation", type=int, default=10)
    a("--gpu_ids", type=str, default="0")
    a("--use_fixed_testset", action="store_true", default=False)
    a("--autoregressive", type=str, default="")
    a("--num_split", type=int, default=1)
    a("--vqvae", action="store_true", default=False)
    a("--load_vqvae", action="store_true", default=False)
    a("--vqvae_path", type=str, default="")
    a("--dataset", type=str, default="")
    # additions (not in the reference): where the input image lives, and the seed of the synthetic weights
    a("--demo_folder", type=str, default="demo", help="directory of --demo_img_name (the reference hard-codes 'demo')")
    a("--seed", type=int, default=0)
    return p


def load_checkpoint(path):
    """-> (state_dict without xyzs/ones buffers, opts Namespace or None) -- demo.py:201-208, opts_helper.py:6."""
    ck = torch.load(path, map_location="cpu", weights_only=False)
    if "state_dict" not in ck:
        raise KeyError("%s: not a PixelSynth checkpoint (no 'state_dict' entry)" % path)
    sd = {k: v for k, v in ck["state_dict"].items() if not ("xyzs" in k) and not ("ones" in k)}
    return sd, ck.get("opts")


def opts_helper(test_ops, ck_opts=None):
    """utils/opts_helper.py:3-55: the checkpoint's training Namespace overridden by the test-time flags."""
    if ck_opts is None:
        opts = types.SimpleNamespace(**DEFAULT_OPTS)
    else:
        opts = types.SimpleNamespace(**vars(ck_opts)) if not isinstance(ck_opts, dict) else types.SimpleNamespace(**ck_opts)
    t = vars(test_ops)
    opts.isTrain = True
    opts.only_high_res = False
    opts.lr_d = 0.001
    for k in ("pretrain", "background_smoothing_kernel_size", "decoder_truncation_threshold", "temperature", "temp_eps",
              "val_rotation", "dataset_folder"):
        setattr(opts, k, t[k])
    if t.get("use_fixed_testset") is not None:
        opts.use_fixed_testset = t["use_fixed_testset"]
    if t.get("vqvae") is not None:
        opts.vqvae = t["vqvae"]
    if t.get("num_split", 0) > 0:
        opts.num_split = t["num_split"]
    for k in ("num_samples", "directions", "direction", "rotation", "model_setting", "demo_img_name",
              "sequential_outpainting"):
        if k in t:
            setattr(opts, k, t[k])
    if "dataset" in t and (t["dataset"] or not hasattr(opts, "dataset")):
        opts.dataset = t["dataset"]   # the reference overwrites unconditionally (:41-42); '' would erase the trained value
    opts.homography = t["homography"]
    opts.no_outpainting = t["no_outpainting"]
    opts.normalize_before_residual = False
    if getattr(opts, "dataset", "") == "test_mp3d" or t["old_model"] == "modelcheckpoints/mp3d/pixelsynth.pth":
        opts.normalize_before_residual = True
    opts.render_ids = t["render_ids"]
    opts.gpu_ids = t["gpu_ids"]
    opts.train_depth = False
    if ck_opts is None or not hasattr(opts, "seed"):
        opts.seed = t.get("seed", 0)
    return opts


def assemble_state(test_ops, state_dict):
    """Applies the --load_vqvae / --load_autoregressive overlays of demo.py:208-221 to the (prefix-stripped) state."""
    from .models.z_buffermodel import strip_parallel_prefixes

    sd = strip_parallel_prefixes(state_dict)
    if test_ops.load_vqvae:
        extra = torch.load(test_ops.vqvae_path, map_location="cpu", weights_only=False)
        sd = {k: v for k, v in sd.items() if not k.startswith("vqvae.")}
        sd.update({"vqvae." + k[7:]: v for k, v in extra.items()})          # k[7:] removes `module.` (:213-214)
    if test_ops.load_autoregressive:
        extra = torch.load(test_ops.autoregressive, map_location="cpu", weights_only=False)["model_state_dict"]
        sd.update({"outpaint2." + k: v for k, v in extra.items()})           # strict=False overlay (:220-221)
    return sd


def process_demo_data(opts, folder="demo"):
    """demo.py:27-98: the input image resized to W x W (PIL bilinear, what torchvision's Resize does to a PIL image),
    scaled to [-1, 1]; cameras P = (offset . origK) . [I|0] with the y-flip / negative-z convention of Habitat, K = I."""
    from PIL import Image

    image = Image.open(os.path.join(folder, opts.demo_img_name))
    shape = np.array(image).shape
    ratio = shape[1] / shape[0]
    W = int(opts.W)
    arr = np.array(image.resize((W, W), Image.BILINEAR))
    if arr.ndim == 2:
        arr = arr[:, :, None]
    x = torch.from_numpy(np.ascontiguousarray(arr)).permute(2, 0, 1).float().div(255)   # ToTensor
    x = (x - 0.5) / 0.5                                                                  # Normalize(.5, .5)
    offset = np.array([[2, 0, -1], [0, -2, 1], [0, 0, -1]], dtype=np.float32)
    K = np.eye(4, dtype=np.float32)
    invK = np.linalg.inv(K)
    extrinsics = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0]])
    intrinsics = np.array([1.0, 1.0 * ratio, 0.5, 0.5])
    origK = np.array([[intrinsics[0], 0, intrinsics[2]], [0, intrinsics[1], intrinsics[3]], [0, 0, 1]], dtype=np.float32)
    P = np.matmul(np.matmul(offset, origK), extrinsics)
    P = np.vstack((P, np.zeros((1, 4), dtype=np.float32))).astype(np.float32)
    P[3, 3] = 1
    Pinv = np.linalg.inv(P)
    cam = {"P": torch.tensor(P).unsqueeze(0), "Pinv": torch.tensor(Pinv).unsqueeze(0),
           "OrigP": torch.tensor(extrinsics).unsqueeze(0), "K": torch.tensor(K).unsqueeze(0),
           "Kinv": torch.tensor(invK).unsqueeze(0)}
    return {"images": [x.unsqueeze(0)], "cameras": [cam]}


def save_image(tensor, path, nrow=8, padding=2):
    """torchvision.utils.save_image for (B,C,H,W) in [0,1]: one image as is, several as a grid with 2 px of padding;
    x*255 + 0.5, clamped, truncated to uint8."""
    from PIL import Image

    t = tensor.detach().float().cpu()
    if t.dim() == 3:
        t = t.unsqueeze(0)
    if t.shape[1] == 1:
        t = t.repeat(1, 3, 1, 1)
    b, c, h, w = t.shape
    if b == 1:
        grid = t[0]
    else:
        xm = min(nrow, b)
        ym = (b + xm - 1) // xm
        grid = torch.zeros(c, (h + padding) * ym + padding, (w + padding) * xm + padding)
        for k in range(b):
            y, x = divmod(k, xm)
            grid[:, padding + y * (h + padding):padding + y * (h + padding) + h,
                 padding + x * (w + padding):padding + x * (w + padding) + w] = t[k]
    arr = grid.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    Image.fromarray(arr).save(path)


def scene_splits(direction, num_split):
    if direction in ("U", "D", "UL", "UR", "DR", "DL"):
        return max(int(num_split / 2), 1)
    return num_split


def save_scene(pred_imgs, test_ops):
    """demo.py:100-121."""
    for direction in test_ops.directions:
        if direction in ("S", "C"):
            continue
        for i in range(1, scene_splits(direction, test_ops.num_split) + 1):
            save_image(pred_imgs["PredImg_%s_%d" % (direction, i)],
                       test_ops.result_folder + "/scene/output_image_%s_%04d.png" % (direction, i))


def video_frames(num_split):
    """Frame order of demo.py:123-161: R out and back, L out and back, two circles, two translation loops."""
    keys = ["PredImg_R_0"]
    for direction in ("R", "L", "C", "C", "S", "S"):
        n = num_split * 2 if direction in ("S", "C") else num_split
        keys += ["PredImg_%s_%d" % (direction, i) for i in range(1, n)]
        if direction not in ("S", "C"):
            keys += ["PredImg_%s_%d" % (direction, i) for i in range(n - 1, -1, -1)]
    return keys


def save_video(pred_imgs, test_ops):
    for ct, k in enumerate(video_frames(test_ops.num_split)):
        save_image(pred_imgs[k], test_ops.result_folder + "/video/%d.png" % ct)


def save_img(pred_imgs, test_ops):
    """demo.py:163-176 (the `%d` of a fractional rotation truncates to 0 there as well)."""
    save_image(pred_imgs["PredImg"], test_ops.result_folder + "/output_image_%s_%d.png" % (test_ops.direction, test_ops.rotation))
    if pred_imgs["FeaturesImg"].shape[1] == 3:
        save_image(pred_imgs["FeaturesImg"],
                   test_ops.result_folder + "/input_fs_image_%s_%d.png" % (test_ops.direction, test_ops.rotation))


def main(argv=None):
    test_ops = build_parser().parse_args(argv)
    state, ck_opts = (None, None)
    if test_ops.old_model:
        state, ck_opts = load_checkpoint(test_ops.old_model)
        state = assemble_state(test_ops, state)
    else:
        print("no --old_model: seeded random weights (seed %d), shipped RealEstate10K options" % test_ops.seed)
    opts = opts_helper(test_ops, ck_opts)
    if not torch.cuda.is_available():
        raise RuntimeError("pixelsynth_b200.demo needs a CUDA device (sm_100a); there is no CPU path")
    device = "cuda:" + str([int(g.strip()) for g in opts.gpu_ids.split(",")][0])   # batch 1: DataParallel used one GPU too
    torch.cuda.set_device(device)
    from .models.base_model import BaseModel
    from .models.z_buffermodel import ZbufferModelPts

    model_to_test = BaseModel(ZbufferModelPts(opts, state_dict=state, device=device), opts)
    print("Loaded models...")
    batch = process_demo_data(opts, test_ops.demo_folder)
    with torch.no_grad():
        _, pred_imgs, _ = model_to_test(batch, isval=True, return_batch=True)
    save_image(pred_imgs["InputImg"], test_ops.result_folder + "/input_image_.png")
    if opts.model_setting == "gen_scene":
        save_scene(pred_imgs, test_ops)
        if all(k in pred_imgs for k in video_frames(test_ops.num_split)):
            save_video(pred_imgs, test_ops)   # the reference's video needs the R, L, C, S sweeps (demo.py:133)
    else:
        save_img(pred_imgs, test_ops)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
