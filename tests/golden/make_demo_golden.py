"""Golden vectors for the demo front end and the target-camera synthesis (run in the build container only; needs
/root/reference).  Nothing is re-typed: the reference's OWN source of
  demo.py:27-98                    process_demo_data      (image transform + demo cameras)
  models/z_buffermodel.py:186-242  eulerAnglesToRotationMatrix, get_rt_from_rot
  models/z_buffermodel.py:264-276  the rank fusion at the end of get_best_sample
is cut out of the files with `ast` and executed here on CPU (`.cuda()` patched to the identity; demo.py and
z_buffermodel.py cannot be imported whole: they need pytorch3d / checkpoints / wget).  Outputs:
  tests/golden/demo_input.png   a small non-square synthetic photograph (the demo's input file)
  tests/golden/demo_front.npz   its transformed tensor + cameras, target cameras per case, rank-fusion cases
"""
import ast
import math
import os
import sys
import types

import numpy as np
import torch
from PIL import Image
from torchvision import transforms as trn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def cut(path, names, cls=None):
    src = open(path).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    return {n.name: n for n in body if isinstance(n, ast.FunctionDef) and n.name in names}, src


def main():
    torch.Tensor.cuda = lambda self, *a, **k: self
    out = {}

    # ---- demo.py: process_demo_data --------------------------------------------------------------
    rng = np.random.default_rng(7)
    yy, xx = np.mgrid[0:80, 0:120]
    img = np.stack([127 + 120 * np.sin(xx / 9.0 + c) * np.cos(yy / 7.0 - c) for c in range(3)], -1)
    img = np.clip(img + rng.normal(0, 6, img.shape), 0, 255).astype(np.uint8)
    png = os.path.join(HERE, "demo_input.png")
    Image.fromarray(img).save(png)
    fns, src = cut(os.path.join(REF, "demo.py"), {"process_demo_data"})
    ns = {"np": np, "torch": torch, "trn": trn, "Image": Image, "os": os}
    exec(compile(ast.Module([fns["process_demo_data"]], []), "demo.py", "exec"), ns)
    tmp = "/tmp/_demo_golden"
    os.makedirs(os.path.join(tmp, "demo"), exist_ok=True)
    Image.fromarray(img).save(os.path.join(tmp, "demo", "demo_input.png"))
    cwd = os.getcwd()
    os.chdir(tmp)
    batch = ns["process_demo_data"](types.SimpleNamespace(W=256, demo_img_name="demo_input.png"))
    os.chdir(cwd)
    out["image"] = batch["images"][0].numpy()
    for k in ("P", "Pinv", "K", "Kinv"):
        out["cam_" + k] = batch["cameras"][0][k].numpy()

    # ---- z_buffermodel.py: get_rt_from_rot ---------------------------------------------------------
    fns, _ = cut(os.path.join(REF, "models/z_buffermodel.py"), {"eulerAnglesToRotationMatrix", "get_rt_from_rot",
                                                                "get_best_sample"}, cls="ZbufferModelPts")
    ns = {"np": np, "torch": torch, "math": math}
    exec(compile(ast.Module([fns["eulerAnglesToRotationMatrix"], fns["get_rt_from_rot"]], []), "z_buffermodel.py", "exec"), ns)
    Shim = type("Shim", (), {"eulerAnglesToRotationMatrix": ns["eulerAnglesToRotationMatrix"],
                             "get_rt_from_rot": ns["get_rt_from_rot"]})
    rotvecs = {'R': np.array([0, .6, 0]), 'L': np.array([0, -.6, 0]), 'U': np.array([-.3, 0, 0]), 'D': np.array([.3, 0, 0]),
               'UR': np.array([-.15, .3, 0]), 'UL': np.array([-.15, -.3, 0]), 'DR': np.array([.15, .3, 0]),
               'DL': np.array([.15, -.3, 0])}    # z_buffermodel.py:112-113 (attribute of the instance, not a function)
    input_RT = batch["cameras"][0]["P"].clone()
    cases = []
    for d in rotvecs:
        for rot in (0.6, 0.3):
            for hom in (False, True):
                cases.append(("gen_img", d, rot, hom, -1, -1))
    for d in list(rotvecs) + ["S", "C"]:
        for num, den in ((0, 4), (1, 4), (3, 4), (4, 4), (5, 64), (37, 64)):
            cases.append(("gen_scene", d, 0.6, False, num, den))
    cases.append(("gen_scene", "R", 0.6, True, 2, 4))
    rts, rtinvs = [], []
    for setting, d, rot, hom, num, den in cases:
        s = Shim()
        s.rotvecs = rotvecs
        s.opt = types.SimpleNamespace(model_setting=setting, rotation=rot, homography=hom)
        inv, rt = s.get_rt_from_rot(d, input_RT.clone(), None if num < 0 else num, None if den < 0 else den)
        rts.append(rt.numpy().astype(np.float32))
        rtinvs.append(inv.numpy().astype(np.float32))
    out["rt_cases"] = np.array(["%s|%s|%r|%d|%d|%d" % c for c in [(a, b, c_, int(h), n, e) for a, b, c_, h, n, e in cases]])
    out["rt"] = np.stack(rts)
    out["rtinv"] = np.stack(rtinvs)

    # ---- z_buffermodel.py: rank fusion (the statements of get_best_sample after its sampling loop) --
    gb = fns["get_best_sample"]
    tail = [st for st in gb.body[2:] if not isinstance(st, ast.Return)]              # everything after the sampling loop
    code = compile(ast.Module(tail, []), "rank", "exec")
    rng = np.random.default_rng(3)
    ds, es, bests = [], [], []
    for n in (2, 3, 5, 8, 8, 50):
        d = rng.normal(0, 1, n).astype(np.float32)
        e = rng.uniform(1, 5, n)
        if n == 8 and len(ds) == 4:
            d[3] = d[5]          # ties
            e[1] = e[2]
        env = {"np": np, "discrim_scores": [torch.tensor(v) for v in d], "entropy_scores": list(e),
               "imgs": list(range(n)), "self": types.SimpleNamespace(opt=types.SimpleNamespace(num_samples=n))}
        exec(code, env)
        ds.append(np.pad(d, (0, 50 - n)))
        es.append(np.pad(e, (0, 50 - n)))
        bests.append([n, int(env["best"])])
    out["rank_d"], out["rank_e"], out["rank_best"] = np.stack(ds), np.stack(es), np.array(bests)

    np.savez_compressed(os.path.join(HERE, "demo_front.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
