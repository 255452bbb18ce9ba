"""Choosing the best of num_samples > 1 outpainting candidates: ZbufferModelPts.get_best_sample
(models/z_buffermodel.py:244-276).

The reference scores every candidate twice -- the discriminator's D_Fake loss on (candidate, input image)
(:254, models/losses/gan_loss.py:172-189) and the entropy of a places365 resnet18's class distribution (:256-261) --
turns each score list into ranks and keeps the candidate with the largest
    0.5 * (n - 1 - entropy_rank) + 0.5 * discriminator_rank                                        (:264-276)
i.e. low scene-classification entropy and high D_Fake.  The rank fusion is reproduced here exactly (host integers;
pinned to the reference's own statements by tests/golden/make_demo_golden.py).  GpuRanker scores ALL candidates in one
batch on the sm_100a kernels (nets.MultiscaleDiscriminatorB200, nets.ResNet18B200) -- two device->host reads per call
instead of the reference's two per candidate; Ranker keeps the injected-callable form.
"""
import numpy as np
import torch


def rank_fusion(discrim_scores, entropy_scores):
    """-> index of the best candidate (z_buffermodel.py:264-276).  np.argsort's default (unstable quicksort -> for
    these sizes an insertion/introsort that is deterministic) is what the reference calls; ties are broken the same
    way because the same call is made on the same values."""
    n = len(discrim_scores)
    if n != len(entropy_scores) or n == 0:
        raise ValueError("rank_fusion: need one discriminator and one entropy score per candidate")
    sorted_disc = np.array([float(s) for s in discrim_scores], dtype=np.float32).argsort()
    sorted_entr = np.array([float(s) for s in entropy_scores]).argsort()
    discrim_ranks = np.array([np.where(sorted_disc == i)[0][0] for i in range(n)])
    entropy_ranks = np.array([np.where(sorted_entr == i)[0][0] for i in range(n)])
    total = .5 * (n - 1 - entropy_ranks) + .5 * discrim_ranks
    return int(np.argmax(total))


def entropy_of_logits(logit):
    """-sum p log p of softmax(logit) over the classes (z_buffermodel.py:259-261), float64 like the numpy sum there."""
    p = torch.softmax(logit.detach().float().cpu().reshape(-1), 0).numpy()
    return float(-np.sum(p * np.log(p)))


def hinge_d_fake(pred_fake):
    """D_Fake of the multiscale hinge discriminator loss (gan_loss.py:80-88 with target_is_real=False,
    for_discriminator=True; :101-116 averages over the discriminators): -mean(min(-x - 1, 0)) of each scale's LAST
    feature map, averaged over scales.  pred_fake: list (scale) of lists (layer outputs) or of tensors."""
    losses = []
    for p in pred_fake:
        x = p[-1] if isinstance(p, (list, tuple)) else p
        losses.append(-torch.mean(torch.min(-x - 1, torch.zeros_like(x))))
    return sum(losses) / len(losses)


class Ranker:
    """ranker(imgs, input_img) -> index, as ZbufferModelPts.get_best_sample uses it.
    discriminator(fake, real) -> D_Fake scalar (netD.run_discriminator_one_step(...)["D_Fake"].mean());
    classifier(img (1,3,256,256) in [-1,1]) -> class logits.  A missing scorer contributes equal scores."""

    def __init__(self, discriminator=None, classifier=None):
        self.discriminator = discriminator
        self.classifier = classifier

    def __call__(self, imgs, input_img):
        d = [float(self.discriminator(im, input_img)) if self.discriminator else 0.0 for im in imgs]
        e = [entropy_of_logits(self.classifier(im[:1])) if self.classifier else 0.0 for im in imgs]   # image 0 only (:256)
        return rank_fusion(d, e)


class GpuRanker:
    """ranker(imgs (n,B,3,S,S), input_img) -> index of the best candidate, every score computed on the GPU in one batch:
    D_Fake of the multiscale PatchGAN on each candidate's B images (gan_loss.py:172-181; the real half of the reference's
    fake|real batch does not enter D_Fake and instance norm is per sample, so it is not run), and the entropy of the
    places365 resnet18 on each candidate's image 0 through the reference's reshape + PIL resize (:256-261)."""

    def __init__(self, discriminator, classifier):
        self.discriminator = discriminator
        self.classifier = classifier
        self.last_scores = None

    def __call__(self, imgs, input_img=None):
        n, B = imgs.shape[:2]
        d = self.discriminator.d_fake(imgs.reshape(n * B, *imgs.shape[2:]), n)
        e = self.classifier.entropy(imgs)
        d, e = d.float().cpu().numpy(), e.double().cpu().numpy()
        self.last_scores = (d, e)
        return rank_fusion(list(d), list(e))
