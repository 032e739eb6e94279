"""Self-critical CIDEr-D reward on the GPU (libisc_b200.so), behind the reference's reward API.

Drop-in for /root/reference/self_critical/utils.py:38-83 (get_ciderd_scorer,
get_self_critical_reward, RewardCriterion :169-177) and for
/root/reference/self_critical/cider/pyciderevalcap/ciderD/ciderD.py:16-48 (CiderD).
Captions are id sequences; the reference turns them into space-joined id strings
(utils._array_to_str) — here they stay integers end to end and the hypotheses never leave the
device. No CPU fallback: scoring needs the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib

MAX_WORDS = 32  # words per scored caption after _array_to_str (kernel limit, incl. the appended EOS)
MAX_WORDS_DF = 64  # words per caption of the document-frequency corpus


def ids_to_words(arr, sos_token, eos_token):
    """utils._array_to_str (utils.py:11-21) on ids: drop a leading SOS, cut at EOS, append EOS."""
    arr = [int(x) for x in arr]
    if arr and arr[0] == sos_token:
        arr = arr[1:]
    out = []
    for w in arr:
        if w == eos_token:
            break
        out.append(w)
    out.append(eos_token)
    return out


class RefSet:
    """References of a set of images, packed on the device: tokens int32 [R, ld], lens int32 [R],
    offsets int32 [N+1]."""

    def __init__(self, refs_words, device, max_words=MAX_WORDS):
        n_refs = sum(len(r) for r in refs_words)
        ld = max([len(w) for r in refs_words for w in r] + [1])
        if ld > max_words:
            raise ValueError("reference captions longer than %d words are not supported (got %d); truncate "
                             "them like the reference's dataloader does (max_seq_len + 1 ids)" % (max_words, ld))
        tok = np.zeros((max(n_refs, 1), ld), dtype=np.int32)
        lens = np.zeros(max(n_refs, 1), dtype=np.int32)
        offs = np.zeros(len(refs_words) + 1, dtype=np.int32)
        r = 0
        for i, refs in enumerate(refs_words):
            for w in refs:
                tok[r, :len(w)] = w
                lens[r] = len(w)
                r += 1
            offs[i + 1] = r
        if tok.size and (tok.min() < 0 or tok.max() >= 65535):
            raise ValueError("token ids must be in [0, 65534] for the packed 16-bit n-gram keys")
        self.n_images = len(refs_words)
        self.ld = ld
        self.n_positions = int(sum(max(4 * int(l) - 6, int(l)) for l in lens))
        # pinned staging + async copies: building the references of a batch must not drain the stream (RL iteration)
        self.tokens = _lib.to_device_async(torch.from_numpy(tok), torch.from_numpy(tok).dtype, device)
        self.lens = _lib.to_device_async(torch.from_numpy(lens), torch.from_numpy(lens).dtype, device)
        self.offsets = _lib.to_device_async(torch.from_numpy(offs), torch.from_numpy(offs).dtype, device)


class CiderD:
    """CIDEr-D scorer with the document-frequency table resident on the GPU.

    ``refs``: corpus for the DF table — list (per image) of list of captions, each caption either a
    space-joined string of ids (the reference's format, ciderD.py:20-21) or a list of ids."""

    def __init__(self, n=4, sigma=6.0, refs=None, device=None):
        if n != 4 or float(sigma) != 6.0:
            raise ValueError("the CUDA CIDEr-D kernel is compiled for n=4, sigma=6.0 (the reference's values)")
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("CiderD runs on the GPU only (no CPU fallback)")
        self.table = None
        self.ref_len = None
        self._registered = None
        self._fn_index = None
        if refs:
            self.update_df(refs)

    @staticmethod
    def _words(cap):
        if isinstance(cap, str):
            return [int(w) for w in cap.split()]
        return [int(w) for w in cap]

    def update_df(self, refs):
        lib = _lib.load()
        words = [[self._words(c) for c in caps] for caps in refs]
        rs = RefSet(words, self.device, MAX_WORDS_DF)
        slots = 1 << max(10, int(math.ceil(math.log2(max(2 * rs.n_positions, 2)))))
        self.slots = slots
        self.table = torch.empty(lib.isc_cider_table_bytes(slots), dtype=torch.uint8, device=self.device)
        flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.isc_cider_build_df(_lib.ptr(rs.tokens), _lib.ptr(rs.lens), rs.ld, _lib.ptr(rs.offsets),
                                              rs.n_images, _lib.ptr(self.table), slots, _lib.ptr(flag),
                                              _lib.stream_ptr(self.device)), "isc_cider_build_df")
        if int(flag.item()) != 0:
            raise RuntimeError("CIDEr DF build overflow (an image with > 2048 n-gram positions or a full table)")
        self.ref_len = math.log(float(len(refs)))  # ciderD_scorer.py:88
        return self

    # ---- tensor API (hypotheses stay on the device) ----
    def register_ground_truth(self, fns, ground_truth, sos_token, eos_token):
        """Pack ``ground_truth[fn]`` (lists of ids) for every fn once; later calls index it by fn."""
        words = [[ids_to_words(c, sos_token, eos_token) for c in ground_truth[fn]] for fn in fns]
        self._registered = RefSet(words, self.device)
        self._fn_index = {fn: i for i, fn in enumerate(fns)}
        return self._registered

    def score_ids(self, hyps, hyp_img, refset, sos_token, eos_token):
        """hyps int64 [N,T] on the device, hyp_img int32 [N] indexes ``refset``. -> float64 [N] (x10)."""
        lib = _lib.load()
        if self.table is None:
            raise RuntimeError("CiderD: document-frequency table not built (pass refs=...)")
        hyps = hyps.to(self.device).long().contiguous()
        hyp_img = _lib.to_device_async(hyp_img, torch.int32, self.device).contiguous()
        N, T = hyps.shape
        scores = torch.empty(N, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.isc_cider_score(_lib.ptr(self.table), self.slots, self.ref_len, _lib.ptr(hyps), T,
                                           _lib.ptr(hyp_img), N, _lib.ptr(refset.tokens), _lib.ptr(refset.lens),
                                           refset.ld, _lib.ptr(refset.offsets), int(sos_token), int(eos_token),
                                           _lib.ptr(scores), _lib.stream_ptr(self.device)), "isc_cider_score")
        return scores

    def ngram_counts(self, hyp, sos_token, eos_token):
        """{packed key: term frequency} of one hypothesis (parity tests: exact integer counts)."""
        lib = _lib.load()
        hyp = hyp.to(self.device).long().contiguous().reshape(-1)
        keys = torch.zeros(128, dtype=torch.int64, device=self.device)
        cnts = torch.zeros(128, dtype=torch.int32, device=self.device)
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.isc_cider_ngram_counts(_lib.ptr(hyp), hyp.numel(), int(sos_token), int(eos_token),
                                                  _lib.ptr(keys), _lib.ptr(cnts), _lib.ptr(n),
                                                  _lib.stream_ptr(self.device)), "isc_cider_ngram_counts")
        n = int(n.item())
        ks = keys[:n].cpu().numpy().astype(np.uint64)
        return {int(k): int(c) for k, c in zip(ks, cnts[:n].tolist())}

    # ---- reference API (ciderD.py:24-48): strings of ids in, (mean, ndarray) out ----
    def compute_score(self, gts, res):
        ids = [r["image_id"] for r in res]
        uniq = list(dict.fromkeys(ids))
        for r in res:
            assert type(r["caption"]) is list and len(r["caption"]) == 1
        for u in uniq:
            assert type(gts[u]) is list and len(gts[u]) > 0
        refset = RefSet([[self._words(c) for c in gts[u]] for u in uniq], self.device)
        index = {u: i for i, u in enumerate(uniq)}
        hyp_words = [self._words(r["caption"][0]) for r in res]
        scores = self._score_word_lists(hyp_words, [index[i] for i in ids], refset).cpu().numpy()
        return float(np.mean(scores)), scores

    def _score_word_lists(self, hyp_words, hyp_img, refset):
        """Score hypotheses given as FINAL word lists (the output of utils._array_to_str, whose last word
        is the EOS it appended). The kernel re-applies "cut at the first EOS, append EOS", which is the
        identity on such lists as long as the final word does not occur earlier in the caption."""
        N = len(hyp_words)
        T = max(len(w) for w in hyp_words)
        if T > MAX_WORDS - 1:
            raise ValueError("hypotheses longer than %d words are not supported" % (MAX_WORDS - 1))
        out = torch.empty(N, dtype=torch.float64, device=self.device)
        img = torch.as_tensor(hyp_img, dtype=torch.int32)
        groups = {}
        for i, w in enumerate(hyp_words):
            if not w or w[-1] in w[:-1]:
                raise ValueError("CiderD.compute_score expects captions produced by _array_to_str (unique "
                                 "trailing EOS); use score_ids() for raw id tensors")
            groups.setdefault(w[-1], []).append(i)
        for eos_val, rows in groups.items():
            arr = np.full((len(rows), T), eos_val, dtype=np.int64)
            for j, i in enumerate(rows):
                arr[j, :len(hyp_words[i])] = hyp_words[i]
            rows_t = torch.as_tensor(rows)
            out[rows_t.to(self.device)] = self.score_ids(torch.from_numpy(arr), img[rows_t], refset, -7, eos_val)
        return out


def get_ciderd_scorer(split_captions, sos_token, eos_token, device=None):
    """utils.get_ciderd_scorer (utils.py:38-53): DF table from every caption of every split."""
    captions = {}
    for caps in split_captions.values():
        captions.update(caps)
    refs = [[ids_to_words(c, sos_token, eos_token) for c in caps] for caps in captions.values()]
    return CiderD(refs=refs, device=device)


def self_critical_scores(sample_captions, greedy_captions, fns, ground_truth, sos_token, eos_token, scorer):
    """Device tensor float64 [2B]: CIDEr-D x10 of the B sampled then the B greedy captions."""
    B = len(fns)
    assert sample_captions.shape[0] == greedy_captions.shape[0] == B
    if scorer._fn_index is not None and all(fn in scorer._fn_index for fn in fns):
        refset = scorer._registered
        idx = [scorer._fn_index[fn] for fn in fns]
    else:
        refset = RefSet([[ids_to_words(c, sos_token, eos_token) for c in ground_truth[fn]] for fn in fns],
                        scorer.device)
        idx = list(range(B))
    hyps = torch.cat([sample_captions, greedy_captions], dim=0)
    img = torch.tensor(idx + idx, dtype=torch.int32)
    return scorer.score_ids(hyps, img, refset, sos_token, eos_token)


def get_self_critical_reward(sample_captions, greedy_captions, fns, ground_truth, sos_token, eos_token, scorer,
                             as_tensor=False):
    """utils.get_self_critical_reward (utils.py:56-83): float64 [B,T] = CIDEr(sample) - CIDEr(greedy),
    repeated over T. Returns a numpy array like the reference, or the device tensor with as_tensor=True."""
    if not isinstance(scorer, CiderD):
        raise Exception("do not support this scorer: %s" % type(scorer))
    lib = _lib.load()
    B, T = sample_captions.shape
    scores = self_critical_scores(sample_captions, greedy_captions, fns, ground_truth, sos_token, eos_token, scorer)
    rewards = torch.empty(B, T, dtype=torch.float64, device=scorer.device)
    with torch.cuda.device(scorer.device):
        _lib.check(lib.isc_self_critical_reward(_lib.ptr(scores), B, T, _lib.ptr(rewards),
                                                _lib.stream_ptr(scorer.device)), "isc_self_critical_reward")
    return rewards if as_tensor else rewards.cpu().numpy()


class RewardCriterion(nn.Module):
    """REINFORCE loss (utils.py:169-177)."""

    def forward(self, seq_logprobs, seq_masks, reward):
        out = -seq_logprobs * seq_masks * reward
        return out.sum() / seq_masks.sum()
