"""Per-stage device time of the full pipeline (CUDA events), for DESIGN.md / bench planning."""
import sys, os, types, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from test_pipeline_gpu import make_opt, make_batch
from pixelsynth_b200.models.z_buffermodel import ZbufferModelPts
from pixelsynth_b200 import lmconv

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

def main():
    m = ZbufferModelPts(make_opt(model_setting="gen_paired_img"))
    for B in (1, 8, 32):
        batch = make_batch(B)
        for it in range(2):
            K, K_inv, RT1, RT1i, RT2, RT2i, img, _ = m.process_batch(batch)
            torch.cuda.synchronize(); t = [ev()]
            depth = m.pts_regressor.forward(img, 0.5, 10.0); t.append(ev())
            gen_fs, bg = m.pts_transformer.forward_justpts(img, depth, K, K_inv, RT1, RT1i, RT2, RT2i); t.append(ev())
            torch.cuda.synchronize(); h0 = time.perf_counter()
            _, order, words, smask = lmconv.glue_host(bg); h1 = time.perf_counter(); t.append(ev())
            codes = m.vqvae.encode_top(gen_fs); t.append(ev())
            u = torch.rand(B, 1024)
            sampled = m.outpaint2.sample(codes, order, words, smask, u, 0.7); t.append(ev())
            ar = m.vqvae.decode_code(sampled); t.append(ev())
            comb = m.get_combined(gen_fs, ar, bg); t.append(ev())
            out = m.projector.forward(comb, bg); t.append(ev())
            torch.cuda.synchronize()
        names = ["unet", "splat", "glue(host)", "vq_encode", "lmconv_sample", "vq_decode", "combine", "decoder"]
        ms = [t[i].elapsed_time(t[i + 1]) for i in range(len(names))]
        print("B=%d  " % B + "  ".join("%s %.2f" % (n, v) for n, v in zip(names, ms)) + "  | total %.2f ms, glue host %.2f ms, sampled cells/img %.0f" % (sum(ms), (h1 - h0) * 1e3, smask.reshape(B, -1).sum(1).mean()))

main()
