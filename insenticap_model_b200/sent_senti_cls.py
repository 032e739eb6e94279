"""Drop-in ``SentenceSentimentClassifier`` (inference) on libisc_b200.so — /root/reference/models/sent_senti_cls.py:6-72.

Same constructor and parameter names (``word_embed.0.weight``, ``rnn.weight_ih_l0`` ..., ``excitation.0/2``,
``sent_senti_cls.0/3``: reference checkpoints load verbatim), ``forward(seqs, lengths) -> (pred, word_weights)`` and
``sample``. eval() semantics only (training the classifier, train_sent_senti_cls_rnn.py, is outside this repo's path).
Used by the RL step's classifier reward and the XE pseudo-labels. No CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class SentenceSentimentClassifier(nn.Module):
    def __init__(self, idx2word, sentiment_categories, settings):
        super().__init__()
        if settings["word_emb_dim"] != 512 or settings["rnn_hid_dim"] != 512:
            raise ValueError("libisc_b200 is compiled for word_emb_dim = rnn_hid_dim = 512")
        self.sentiment_categories = sentiment_categories
        self.pad_id = idx2word.index("<PAD>")
        self.vocab_size = len(idx2word)
        self.word_embed = nn.Sequential(nn.Embedding(self.vocab_size, 512, padding_idx=self.pad_id), nn.ReLU(),
                                        nn.Dropout(settings["dropout_p"]))
        self.rnn = nn.LSTM(512, 512, bidirectional=False)
        self.drop = nn.Dropout(settings["dropout_p"])
        self.excitation = nn.Sequential(nn.Linear(512, 512), nn.ReLU(), nn.Linear(512, 512), nn.Sigmoid())
        self.squeeze = nn.AdaptiveAvgPool1d(1)
        self.sent_senti_cls = nn.Sequential(nn.Linear(512, 512), nn.ReLU(), nn.Dropout(settings["dropout_p"]),
                                            nn.Linear(512, len(sentiment_categories)))
        self._packed = None
        self._packed_key = None
        self._ws = None

    def _pack(self, dev):
        lib = _lib.load()
        ps = [self.word_embed[0].weight, self.rnn.weight_ih_l0, self.rnn.weight_hh_l0, self.rnn.bias_ih_l0,
              self.rnn.bias_hh_l0, self.excitation[0].weight, self.excitation[0].bias, self.excitation[2].weight,
              self.excitation[2].bias, self.sent_senti_cls[0].weight, self.sent_senti_cls[0].bias,
              self.sent_senti_cls[3].weight, self.sent_senti_cls[3].bias]
        key = (dev.index,) + tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        n = len(self.sentiment_categories)
        nbytes = lib.isc_sentcls_packed_bytes(self.vocab_size, n)
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        keep = [p.detach().float().contiguous() for p in ps]
        with torch.cuda.device(dev):
            _lib.check(lib.isc_sentcls_pack(self.vocab_size, n, *[_lib.ptr(t) for t in keep], _lib.ptr(packed), nbytes,
                                            _lib.stream_ptr(dev)), "isc_sentcls_pack")
        self._packed, self._packed_key = packed, key
        return packed

    def forward(self, seqs, lengths):
        """seqs int64 [bs, max_seq_len], lengths list/tensor [bs] -> (pred [bs, n], word weights [bs, max(lengths)])."""
        lens = [int(x) for x in (lengths.tolist() if torch.is_tensor(lengths) else lengths)]
        if min(lens) < 1:
            raise RuntimeError("lengths must be >= 1 (pack_padded_sequence rejects empty sequences)")
        T = max(lens)
        if T > seqs.shape[1]:
            raise RuntimeError("a length exceeds the caption tensor's width")
        return self._run(seqs, torch.tensor(lens, dtype=torch.int32), T)

    def forward_device(self, seqs, lengths):
        """The same forward with the lengths as a DEVICE int tensor [bs] (each >= 1) and no host round trip: the word
        weights come back [bs, seqs.shape[1]], with the columns at and past max(lengths) — which the reference's output
        does not have, and which its callers zero-pad (self_critical/utils.py:142-143) — set to zero."""
        T = seqs.shape[1]
        lens = lengths.to(torch.int32)
        pred, att = self._run(seqs, lens, T)
        return pred, att * (torch.arange(T, device=att.device) < lens.max()).to(att.dtype)

    def _run(self, seqs, lens_t, T):
        if self.training:
            raise NotImplementedError("SentenceSentimentClassifier runs in eval() mode on the B200 path")
        dev = self.sent_senti_cls[0].weight.device
        if dev.type != "cuda" or not seqs.is_cuda:
            raise RuntimeError("SentenceSentimentClassifier only exists as sm_100a CUDA kernels: move the module and the "
                               "captions to a B200; there is no CPU fallback")
        lib = _lib.load()
        packed = self._pack(dev)
        B = seqs.shape[0]
        seqs = seqs.long().contiguous()
        lens_t = _lib.to_device_async(lens_t, torch.int32, dev).contiguous()
        n = len(self.sentiment_categories)
        nbytes = lib.isc_sentcls_workspace_bytes(B, T)
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        pred = torch.empty(B, n, dtype=torch.float32, device=dev)
        att = torch.empty(B, T, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.isc_sentcls_forward(self.vocab_size, n, _lib.ptr(packed), _lib.ptr(seqs), seqs.shape[1],
                                               _lib.ptr(lens_t), B, T, _lib.ptr(pred), _lib.ptr(att), _lib.ptr(self._ws),
                                               self._ws.numel(), _lib.stream_ptr(dev)), "isc_sentcls_forward")
        return pred, att

    def sample(self, seqs, lengths):
        self.eval()
        pred, att_weights = self.forward(seqs, lengths)
        result = pred.argmax(-1).tolist()
        return result, [self.sentiment_categories[r] for r in result], att_weights

    def get_optim_and_crit(self, lr, weight_decay=0):
        return torch.optim.Adam(self.parameters(), lr=lr, weight_decay=weight_decay), nn.CrossEntropyLoss()
