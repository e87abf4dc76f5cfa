#!/usr/bin/env python3
"""Converged (4096 spp) oracle renders of the Marschner scenes for north_star's third check (run from the repo root:
python tests/golden/make_converged.py).

BASELINE.json: "converged images (4096 spp) within relMSE < 1e-3 of the reference, plus a per-pixel z-test on the variance
estimate".  The oracle needs minutes for 4096 spp even at 64x64, so its side of the comparison is a committed fixture: per scene
the mean image over 16 batches of 256 sample indices (seed 977) and the per-pixel variance of that mean, estimated from the batch
means.  tests/test_gpu_parity.py::test_converged_marschner_images renders the same scenes on the GPU with an INDEPENDENT seed.
"""
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import orc
import cudapath

# (scene, strand scale, overrides): Marschner hair (hair-curl: four shapes / beckmann; curly-hair: ggx) at reduced strand count
CASES = [('hair-curl', 0.01, dict(width=64, height=64, spp=4096, maxDepth=65)),
         ('curly-hair', 0.01, dict(width=64, height=64, spp=4096, maxDepth=65))]
BATCHES = 16
SEED = 977


def batch_stats(render, spp, seed, develop):
    per = spp // BATCHES
    imgs = []
    for k in range(BATCHES):
        f = render(spp, seed=seed, sample_begin=per * k, sample_end=per * (k + 1))
        imgs.append(develop(f).astype(np.float64))
    b = np.stack(imgs)                                   # (BATCHES, h, w, 3)
    return b.mean(axis=0), b.var(axis=0, ddof=1) / BATCHES


def main():
    out = {}
    for name, scale, ov in CASES:
        env = cudapath.bake_sunsky(**cudapath.scenes.sunsky_params(name))
        osc = orc.scene_from_description(name, scale=scale, overrides=ov, envmap=env)
        t = time.time()
        mean, var = batch_stats(osc.render, ov['spp'], SEED, cudapath.develop)
        print('%s: %d x %d at %d spp in %.1f s, mean %.4f, relative noise %.2e' % (name, ov['width'], ov['height'], ov['spp'], time.time() - t, mean.mean(),
                                                                                    float(np.mean(var / (mean ** 2 + 1e-2)))), flush=True)
        key = name.replace('-', '_')
        out[key + '_mean'] = mean.astype(np.float32); out[key + '_var'] = var.astype(np.float32)
        out[key + '_cfg'] = np.array([scale, ov['width'], ov['height'], ov['spp'], ov['maxDepth'], BATCHES, SEED], np.float64)
    np.savez_compressed(os.path.join(HERE, 'converged_golden.npz'), **out)


if __name__ == '__main__':
    main()
