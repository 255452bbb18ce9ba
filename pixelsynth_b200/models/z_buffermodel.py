"""Mirror of reference models/z_buffermodel.py:29-419 (ZbufferModelPts), inference path only.

Same constructor / forward(batch, netD=None) contract and output keys; every callee on the hot path dispatches to
the sm_100a kernels of libpixelsynth_b200.so:
  pts_regressor  -> nets.UnetB200                 (z_buffermodel.py:41-42, 304-308)
  pts_transformer-> models.projection.PtsManipulator (splat kernels)            (:49-52, 323-332)
  get_masks_for_batch -> lmconv.glue_host (native host code)                    (:641-701)
  vqvae          -> nets.VQVAETopB200                                            (:82, 345, 250)
  sample(outpaint2, ...) -> lmconv.LmconvB200.sample (activation-cached kernel) (:62-74, 246-247)
  get_combined   -> ps_combine                                                   (:703-708)
  projector      -> nets.ResNetDecoderB200                                       (:87, 252)
Weights: a state dict with the reference's key names (`pts_regressor.*`, `vqvae.*`, `outpaint2.*`, `projector.*`;
`module.` / `model.module.` prefixes of DataParallel checkpoints are stripped, demo.py:202-229).  Without a state
dict the networks are random-initialised from opt.seed (pixelsynth_b200/synthetic.py): no checkpoint of the
reference is reachable offline.
num_samples > 1 (z_buffermodel.py:244-276): all candidates are ONE batch (one sampler launch, one decode, one
refinement pass) and `ranker(imgs, input_img) -> index` picks one.  The default ranker is built lazily from the state
dict's `netD.` / `classifier.` entries (seeded random when absent): pixelsynth_b200.ranking.GpuRanker =
nets.MultiscaleDiscriminatorB200 (D_Fake) + nets.ResNet18B200 (places365 entropy) + the reference's rank fusion.
forward_scene (z_buffermodel.py:421-592) takes batches of images (BASELINE config 5).
`sampler_stream` (optional) moves the sampler launch to a side stream: pixelsynth_b200.pipeline.ViewPipeline.
Stochastic elements are explicit: `noise` (decoder, 16 x (B,20)) and `uniforms` (sampler) can be injected; when
absent they are drawn from torch generators seeded as the reference seeds its sampler (sample.py:14-16)."""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, lmconv, nets, synthetic
from .._lib import check
from . import _ops_loaded  # noqa: F401
from .projection.z_buffer_manipulator import PtsManipulator

ROTVECS = {"R": np.array([0, .6, 0]), "L": np.array([0, -.6, 0]), "U": np.array([-.3, 0, 0]), "D": np.array([.3, 0, 0]),
           "UR": np.array([-.15, .3, 0]), "UL": np.array([-.15, -.3, 0]), "DR": np.array([.15, .3, 0]),
           "DL": np.array([.15, -.3, 0])}  # z_buffermodel.py:112-113


def _get(opt, name, default=None):
    return getattr(opt, name, default) if not isinstance(opt, dict) else opt.get(name, default)


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def strip_parallel_prefixes(sd):
    out = {}
    for k, v in sd.items():
        for p in ("model.module.", "module.", "model."):
            if k.startswith(p):
                k = k[len(p):]
                break
        out[k] = v
    return out


def euler_to_matrix(theta):
    """z_buffermodel.py:186-200: R = Rz . Ry . Rx, float64."""
    cx, sx, cy, sy, cz, sz = math.cos(theta[0]), math.sin(theta[0]), math.cos(theta[1]), math.sin(theta[1]), \
        math.cos(theta[2]), math.sin(theta[2])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rz @ (ry @ rx)


class ZbufferModelPts(nn.Module):
    def __init__(self, opt, state_dict=None, device="cuda"):
        super().__init__()
        self.opt = opt
        self.device = device
        seed = int(_get(opt, "seed", 0) or 0)
        if state_dict is None:
            parts = {"pts_regressor.": synthetic.make_state("unet", seed), "vqvae.": synthetic.make_state("vqvae", seed),
                     "outpaint2.": synthetic.make_state("lmconv", seed), "projector.": synthetic.make_state("decoder", seed)}
            state_dict = {p + k: v for p, sd in parts.items() for k, v in sd.items()}
        sd = strip_parallel_prefixes(state_dict)
        self._state, self._seed = sd, seed
        self.pts_regressor = nets.UnetB200(_sub(sd, "pts_regressor."), device)
        self.pts_transformer = PtsManipulator(_get(opt, "W", 256), C=3, opt=opt).to(device)
        self.vqvae = nets.VQVAETopB200(_sub(sd, "vqvae."), device)
        self.outpaint2 = lmconv.LmconvB200(_sub(sd, "outpaint2."), device)
        self.projector = nets.ResNetDecoderB200(_sub(sd, "projector."), device,
                                                bool(_get(opt, "normalize_before_residual", False)))
        self.obs = [3, 32, 32]
        self.ranker = None
        self._bg_pin = None
        self.sampler_stream = None   # optional side stream for the sampler launch (pipelined serving)
        self.front_stream = None     # optional (high-priority) stream for everything up to the VQ-VAE encoder

    # -- z_buffermodel.py:120-184 ---------------------------------------------------------------
    def process_batch(self, batch):
        cam = batch["cameras"][0]
        dev = self.device
        f = lambda t: t.to(dev, non_blocking=t.is_cuda or t.is_pinned()).float()  # pageable sources: blocking copy
        out = [f(cam["K"]), f(cam["Kinv"]), f(cam["P"]), f(cam["Pinv"])]
        if _get(self.opt, "model_setting") in ("train", "gen_paired_img"):
            out += [f(batch["cameras"][-1]["P"]), f(batch["cameras"][-1]["Pinv"]), f(batch["images"][0]),
                    f(batch["images"][-1])]
        else:
            out += [f(batch["images"][0])]
        return out

    # -- z_buffermodel.py:202-242 ---------------------------------------------------------------
    def get_rt_from_rot(self, direction, input_RT, num=None, denom=None):
        num = 0 if num is None else num
        setting = _get(self.opt, "model_setting")
        if setting in ("gen_two_imgs", "gen_scene"):
            if direction == "S":
                new_rt = torch.zeros_like(input_RT)
                new_rt[:, :, :3] = input_RT[:, :, :3]
                new_rt[:, 3, 3] = 1
                # float64 offsets: the float32 translation is promoted, added in double and rounded once (:214)
                t = .35 * torch.tensor([np.sin(2 * np.pi * num / denom), np.cos(2 * np.pi * num / denom),
                                        .4 * np.sin(2 * np.pi * (.25 + num / denom))], device=input_RT.device)
                new_rt[:, :3, 3] = (input_RT[:, :3, 3] + t).to(new_rt.dtype)
                return torch.inverse(new_rt), new_rt
            if direction == "C":
                rotvec = np.array([0.2 * np.cos(2 * np.pi * num / denom), 0.2 * np.sin(2 * np.pi * num / denom), 0])
            else:
                rotvec = ROTVECS[direction] * num / denom
        else:
            rotvec = ROTVECS[direction] * _get(self.opt, "rotation") / np.linalg.norm(ROTVECS[direction])
        mtx = torch.zeros(1, 4, 4, device=input_RT.device)
        mtx[0, 3, 3] = 1
        mtx[0, :3, :3] = torch.tensor(euler_to_matrix(rotvec)).to(torch.float32)
        if _get(self.opt, "homography", False) and direction not in ("S", "C"):
            new_rt = torch.zeros_like(input_RT)
            new_rt[:, :, 3] = input_RT[:, :, 3]
            new_rt[:, :3, :3] = mtx[:, :3, :3] @ input_RT[:, :3, :3]
        else:
            new_rt = mtx @ input_RT
        return torch.inverse(new_rt), new_rt

    def get_masks_for_batch(self, output_RT, input_RTinv, background_mask):
        """-> (dist, order, words, sample_mask) numpy arrays; the reference's three float mask tensors of shape
        (B*513|160|80, 9, 1024) are the nine-bit words[:, 0|1|2] here."""
        return lmconv.glue_host(background_mask)

    def get_combined(self, gen_fs, ar_sample, background_mask):
        b, c, h, w = gen_fs.shape
        out = torch.empty_like(gen_fs)
        m = background_mask.contiguous().view(torch.uint8)
        check(_lib.lib().ps_combine(gen_fs.contiguous().data_ptr(), ar_sample.contiguous().data_ptr(), m.data_ptr(), b, c,
                                    h * w, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "ps_combine")
        return out

    def _sampler_uniforms(self, seed, B):
        # sample.py:14-16 seeds numpy and torch from the sample index; the categorical draws themselves are explicit
        # uniforms here (torch.multinomial's CUDA stream cannot be reproduced by another kernel)
        np.random.seed(seed)
        g = torch.Generator().manual_seed(seed * 10 + int(np.random.randint(188)))
        return torch.rand(B, 1024, generator=g)

    def get_best_sample(self, order, words, sample_mask, codes, background_mask, gen_fs, netD, input_img, noise=None,
                        uniforms=None, prepared=None):
        """z_buffermodel.py:244-276.  num_samples candidates are ONE sampler launch, one decode and one refinement pass
        over a batch of num_samples * B images (candidate i = rows [i*B, (i+1)*B)); the reference loops over them.
        uniforms: (B,n) shared by all candidates, or (num_samples,B,n); None = seeded per candidate like sample.py:14-16.
        noise: (16,B,20) shared, (16,num_samples*B,20), or None = fresh draws like LinearNoiseLayer."""
        n = int(_get(self.opt, "num_samples", 1) or 1)
        B = codes.shape[0]
        T = float(_get(self.opt, "temperature", 1.0))
        if n == 1:
            u = uniforms if uniforms is not None else self._sampler_uniforms(0, B)
            sampled = self._sample(codes, order, words, sample_mask, u, T, prepared)
            ar_sample = self.vqvae.decode_code(sampled)
            combined = self.get_combined(gen_fs, ar_sample, background_mask)
            return self.projector.forward(combined, background_mask, noise)
        if uniforms is None:
            u = torch.cat([self._sampler_uniforms(i, B) for i in range(n)])
        else:
            u = torch.as_tensor(uniforms)
            u = u.reshape(n * B, -1) if u.dim() == 3 else u.repeat(n, 1)
        rep = lambda a: np.concatenate([np.asarray(a)] * n, 0)
        sampled = self.outpaint2.sample(codes.repeat(n, 1, 1), rep(order), rep(words), rep(sample_mask), u, T)
        ar_sample = self.vqvae.decode_code(sampled)
        bg_n = background_mask.repeat(n, 1, 1)
        combined = self.get_combined(gen_fs.repeat(n, 1, 1, 1), ar_sample, bg_n)
        if noise is not None and noise.shape[1] == B:
            noise = noise.repeat(1, n, 1)
        imgs = self.projector.forward(combined, bg_n, noise).view(n, B, *gen_fs.shape[1:])
        if self.ranker is None:
            self.ranker = self._default_ranker()
        best = int(self.ranker(imgs, input_img)) if self.ranker is not None else 0
        self.last_best = best
        return imgs[best]

    def _sample(self, codes, order, words, sample_mask, u, T, prepared=None):
        side = self.sampler_stream
        if side is None:
            return self.outpaint2.sample(codes, order, words, sample_mask, u, T, prepared=prepared)
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            out = self.outpaint2.sample(codes, order, words, sample_mask, u, T, prepared=prepared)
        cur.wait_stream(side)
        out.record_stream(cur)
        return out

    def _default_ranker(self):
        """The reference ranks with BaseModel.netD and ZbufferModelPts.classifier (demo.py:233-243 loads its places365
        weights).  Built lazily from the state dict's `netD.` / `classifier.` entries, or seeded random when absent."""
        from .. import ranking
        sd = self._state
        d = _sub(sd, "netD.netD.") or _sub(sd, "netD.") or synthetic.make_state("netD", self._seed)
        c = _sub(sd, "classifier.") or synthetic.make_state("resnet18", self._seed)
        return ranking.GpuRanker(nets.MultiscaleDiscriminatorB200(d, self.device), nets.ResNet18B200(c, self.device))

    # -- z_buffermodel.py:291-419 ---------------------------------------------------------------
    def forward_image(self, batch, netD=None, noise=None, uniforms=None):
        setting = _get(self.opt, "model_setting")
        if setting == "train":
            raise NotImplementedError("training is out of scope (SURVEY.md section 2)")
        if _get(self.opt, "use_gt_depth", False):
            # the reference reads an undefined `depth_img` on this branch (z_buffermodel.py:318; its assignment at :164 is
            # inside a commented-out block), i.e. it raises NameError there
            raise NotImplementedError("use_gt_depth: dead in the reference too (depth_img is never assigned)")
        if not _get(self.opt, "use_rgb_features", True):
            raise NotImplementedError("the shipped configuration splats RGB (use_rgb_features)")
        # The front end -- everything the host has to wait for before it can build the generation order -- optionally on
        # its own (high-priority) stream; the rest of the step follows on the caller's stream.
        cur = torch.cuda.current_stream()
        front = self.front_stream if self.front_stream is not None else cur
        if front is not cur:
            front.wait_stream(cur)
        with torch.cuda.stream(front):
            if setting == "gen_paired_img":
                K, K_inv, input_RT, input_RTinv, output_RT, output_RTinv, input_img, output_img = self.process_batch(batch)
            else:
                K, K_inv, input_RT, input_RTinv, input_img = self.process_batch(batch)
                output_RTinv, output_RT = self.get_rt_from_rot(_get(self.opt, "direction"), input_RT)
                output_img = None
            if _get(self.opt, "use_inverse_depth", False):
                # (:311-315) 1 / (sigmoid(d) * 10 + 0.01): the affine sigmoid is the last convolution's epilogue, the
                # reciprocal one elementwise pass over the (B,1,S,S) map
                regressed_pts = torch.reciprocal(self.pts_regressor.forward(input_img, 0.01, 10.01))
            else:
                min_z, max_z = float(_get(self.opt, "min_z")), float(_get(self.opt, "max_z"))
                regressed_pts = self.pts_regressor.forward(input_img, min_z, max_z)
            B = input_img.shape[0]
            if output_RT.shape[0] != B:
                output_RT, output_RTinv = output_RT.expand(B, 4, 4), output_RTinv.expand(B, 4, 4)
            gen_fs, background_mask = self.pts_transformer.forward_justpts(
                input_img, regressed_pts, K, K_inv, input_RT, input_RTinv, output_RT.contiguous(), output_RTinv.contiguous())
        if _get(self.opt, "no_outpainting", False):
            raise NotImplementedError("no_outpainting (3-channel decoder, SynSin baseline) is not the shipped configuration")
        else:
            # The generation order / masks / dependency levels are host work on the background mask (native code,
            # csrc/glue.cu + ps_lmconv_levels_host).  The mask goes to pinned memory behind an event, the VQ-VAE
            # encoder is queued, and the host part runs while the GPU encodes.
            if self._bg_pin is None or self._bg_pin.shape != background_mask.shape:
                self._bg_pin = torch.empty(background_mask.shape, dtype=torch.uint8).pin_memory()
            with torch.cuda.stream(front):
                self._bg_pin.copy_(background_mask.view(torch.uint8), non_blocking=True)
                ready = torch.cuda.Event()
                ready.record()
                codes = self.vqvae.encode_top(gen_fs)
            if front is not cur:
                cur.wait_stream(front)
            ready.synchronize()
            _, order, words, sample_mask = self.get_masks_for_batch(output_RT, input_RTinv, self._bg_pin.numpy())
            prepared = self.outpaint2.prepare(order, words, sample_mask, 0)   # passed explicitly: never reused by a later call
            gen_img = self.get_best_sample(order, words, sample_mask, codes, background_mask, gen_fs, netD, input_img,
                                           noise, uniforms, prepared=prepared)
            self.last = dict(depth=regressed_pts, gen_fs=gen_fs, background_mask=background_mask, order=order, words=words,
                             sample_mask=sample_mask, codes=codes, output_RT=output_RT)
        outputs = {"InputImg": input_img, "PredImg": gen_img, "PredDepthImg": regressed_pts / 5 - 1,
                   # (:389) `(~bg).repeat(B,1,1,1).float()`: shape (B,B,S,S), [i,j] = mask j.  Same shape and values as a
                   # broadcast view: at batch 128 the materialised tensor would be 4.3 GB written per step
                   "ForegroundImg": (~background_mask).float().unsqueeze(0).expand(B, -1, -1, -1)}
        if output_img is not None:
            outputs["OutputImg"] = output_img
        outputs["FeaturesImg"] = gen_fs
        loss = {}  # the reference computes loss_function(gen_img, gen_img) here and discards it (:412-414)
        return loss, outputs

    # -- z_buffermodel.py:421-592 ---------------------------------------------------------------
    def _splits(self, direction):
        n = int(_get(self.opt, "num_split", 1))
        if direction in ("S", "C"):
            return n * 2
        if direction in ("U", "D", "UL", "UR", "DR", "DL"):
            return max(n // 2, 1)
        return n

    def _scene_views(self, direction, num_split, sequential):
        """(numerator of the target pose, kind) in rendering order.  Non-sequential outpainting jumps to the far pose
        first ('far': the source camera is the input view, or the pose the previous direction ended on) and then
        walks back through the poses in between; sequential outpainting walks outwards from the input view."""
        if sequential:
            return [(i, "first" if i == 0 else "step") for i in range(num_split + 1)]
        return [(num_split, "far")] + [(i, "step") for i in reversed(range(num_split))]

    def forward_scene(self, batch, netD=None, noise=None, uniforms=None):
        """gen_scene: a sweep of views per direction over ONE growing point cloud.  Every rendered view is fed back as
        the next source image; only the pixels a view had to outpaint are appended to the cloud
        (PtsManipulator.forward_justpts_cumulative).  The reference renders one image at a time (its (1,4,4) pose
        matrices are bmm'ed with the batch); here a batch of B images sweeps in lock step
        (BASELINE config 5): every image gets the same relative camera motion applied to its OWN input pose, the clouds
        are ragged per image (zero-padded, see PtsManipulator._compact), and image b of the batch equals what a batch-1
        call on image b returns."""
        K, K_inv, input_RT, input_RTinv, input_img = self.process_batch(batch)
        if _get(self.opt, "no_outpainting", False):
            raise NotImplementedError("no_outpainting (3-channel decoder, SynSin baseline) is not the shipped configuration")
        min_z, max_z = float(_get(self.opt, "min_z")), float(_get(self.opt, "max_z"))
        sequential = bool(_get(self.opt, "sequential_outpainting", False))
        outputs = {"InputImg": input_img}
        current_img, cloud, feats, last_bg, last_out_inv = input_img, None, None, None, None
        last_num, last_dir = None, None
        self.last_scene = []
        gen_img = input_img
        for direction in _get(self.opt, "directions"):
            num_split = self._splits(direction)
            for num, kind in self._scene_views(direction, num_split, sequential):
                if kind in ("far", "first"):
                    if last_num is not None:
                        src_inv, src_rt = self.get_rt_from_rot(last_dir, input_RT, last_num, num_split)
                    else:
                        src_inv, src_rt = input_RTinv, input_RT
                else:
                    src_inv, src_rt = self.get_rt_from_rot(direction, input_RT, last_num, num_split)
                dst_inv, dst_rt = self.get_rt_from_rot(direction, input_RT, num, num_split)
                depth = self.pts_regressor.forward(current_img, min_z, max_z)
                gen_fs, bg, new_cloud, new_feats = self.pts_transformer.forward_justpts_cumulative(
                    current_img, depth, K, K_inv, src_rt.contiguous(), src_inv.contiguous(), dst_rt.contiguous(),
                    dst_inv.contiguous(), cloud, feats, last_bg, last_out_inv)
                _, order, words, sample_mask = self.get_masks_for_batch(dst_rt, src_inv, bg)
                codes = self.vqvae.encode_top(gen_fs)
                gen_img = self.get_best_sample(order, words, sample_mask, codes, bg, gen_fs, netD, input_img, noise, uniforms)
                self.last_scene.append(dict(direction=direction, num=num, src=current_img, depth=depth, src_rt=src_rt,
                                            src_inv=src_inv, dst_rt=dst_rt, dst_inv=dst_inv, prior_cloud=cloud,
                                            prior_feats=feats, prior_bg=last_bg, prior_out_inv=last_out_inv, gen_fs=gen_fs,
                                            background_mask=bg, cloud=new_cloud))
                tag = "%s_%d" % (direction, num)
                outputs["PredImg_" + tag] = gen_img
                outputs["FeaturesImg_" + tag] = gen_fs
                if kind == "far" or (sequential and num == num_split):
                    outputs["PredDepthImg_" + tag] = depth
                    nb = input_img.shape[0]   # the reference's `(~bg).repeat(B,1,1,1)` quirk, as a broadcast view for B > 1
                    outputs["ForegroundImg_" + tag] = (~bg).repeat(nb, 1, 1, 1).float() if nb == 1 else \
                        (~bg).float().unsqueeze(0).expand(nb, -1, -1, -1)
                    last_dir = direction
                current_img, cloud, feats, last_bg, last_out_inv, last_num = gen_img, new_cloud, new_feats, bg, dst_inv, num
        return {}, outputs

    def forward(self, batch, netD=None, **kw):
        setting = _get(self.opt, "model_setting")
        if setting == "gen_scene":
            return self.forward_scene(batch, netD, **kw)
        if setting == "gen_two_imgs":
            raise NotImplementedError("gen_two_imgs (the consistency evaluation's two-view mode) is evaluation code, out of scope")
        return self.forward_image(batch, netD, **kw)
