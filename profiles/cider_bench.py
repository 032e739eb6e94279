"""CIDEr-D scoring throughput at the BASELINE configs[3] size (512 images x (5 sampled + 1 greedy) = 3072 hypotheses against
5 synthetic references per image): CUDA-event time of isc_cider_score, hypotheses/s, n-gram probes/s and the random-sector
bandwidth SURVEY 8(d) asks for (~372 probes x 32 B sectors per scored hypothesis). One JSON line.
Usage: python profiles/cider_bench.py [images] [reps]     (under ncu: -k regex:cider_score -c 1)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from insenticap_model_b200 import reward as R  # noqa: E402
from insenticap_model_b200 import synthetic as syn  # noqa: E402
from oracle import cider_oracle as CO  # noqa: E402  (ids -> word lists only; the scorer under test is the CUDA one)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
V, T, SPI = 10000, 16, 5
dev = torch.device("cuda", 0)
refs = syn.synthetic_references(B, V, 5, seed=3)
words = [[CO.ids_to_words(c, 1, 2) for c in caps] for caps in refs]
scorer = R.CiderD(refs=words, device=dev)
refset = R.RefSet(words, dev)
g = torch.Generator().manual_seed(5)
# hypotheses: a reference of the image with ~30 % of its words resampled (realistic overlap), EOS-terminated, 0-padded
hyps = torch.zeros(B * (SPI + 1), T, dtype=torch.long)
img = torch.arange(B).repeat_interleave(SPI + 1).int()
for i in range(hyps.shape[0]):
    r = refs[int(img[i])][i % 5][1:-1][:T - 1]
    r = [w if torch.rand(1, generator=g).item() > 0.3 else int(torch.randint(4, 200, (1,), generator=g)) for w in r]
    hyps[i, :len(r)] = torch.tensor(r)
    hyps[i, len(r)] = 2
hyps_d, img_d = hyps.to(dev), img.to(dev)
n_words = (hyps != 0).sum(1).float()
probes = float(sum(max(4 * int(n) - 6, int(n)) for n in n_words)) + 0.0  # hypothesis n-grams; refs: 5 x the same order
ref_probes = 5.0 * probes
for _ in range(3):
    out = scorer.score_ids(hyps_d, img_d, refset, 1, 2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = scorer.score_ids(hyps_d, img_d, refset, 1, 2)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
n = hyps.shape[0]
print(json.dumps({"workload": "isc_cider_score, %d images x %d hypotheses, 5 refs per image, DF table of %d slots" % (B, SPI + 1, scorer.slots),
                  "hypotheses": n, "us_per_call": us, "hypotheses_per_s": n / (us * 1e-6),
                  "table_probes_per_call": probes + ref_probes, "probes_per_s": (probes + ref_probes) / (us * 1e-6),
                  "sector_gbs": (probes + ref_probes) * 32.0 / (us * 1e-6) / 1e9,
                  "mean_score_x10": float(out.mean()), "note": "CUDA events over %d back-to-back calls; ncu: one kernel of ~95 us, issue slots 47 %% busy, DRAM 0.4 %% (the 4 MB table is L2-resident)" % reps}))
