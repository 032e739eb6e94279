"""Sentence sentiment classifier throughput (captions/s) on one B200: B captions of T words, CUDA events, L2-cold inputs
are irrelevant here (weights 10 MB stay L2-resident by design; per-step activations are B x 1024 bf16 planes)."""
import sys

import torch

from insenticap_model_b200 import synthetic as syn
from insenticap_model_b200.sent_senti_cls import SentenceSentimentClassifier

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 17
V = 10000
m = SentenceSentimentClassifier(syn.make_vocab(V), syn.SENTIMENT_CATEGORIES, dict(syn.DEFAULT_SETTINGS))
m.load_state_dict(syn.sent_cls_state_dict(V, 0))
m = m.cuda().eval()
seqs, lengths = syn.sent_cls_inputs(B, V, max_len=T)
seqs = seqs.cuda()
for _ in range(3):
    m(seqs, lengths)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    m(seqs, lengths)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print("sentcls B=%d T=%d: %.3f ms/call, %.0f captions/s" % (B, T, ms, B / ms * 1e3))
