"""Drop-in ``SentimentDetector`` (inference) on libisc_b200.so — /root/reference/models/sentiment_detector.py:5-64.

Same constructor, parameter names (``convs.conv_0.weight`` ... ``senti_conv.weight``, ``output.0.weight`` ...: reference
checkpoints load verbatim), ``forward(features) -> (output, senti_features)`` and ``sample(features, threshold)``. The two
3x3 convolutions run as tcgen05 GEMMs in split-bf16 (isc_senti_detect). eval() semantics only: training the detector
(train_senti.py) is outside this repo's path, so ``forward`` in train() mode raises. No CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class SentimentDetector(nn.Module):
    def __init__(self, sentiment_categories, settings):
        super().__init__()
        self.sentiment_categories = sentiment_categories
        self.neu_idx = sentiment_categories.index("neutral")
        n_convs = settings.get("sentiment_convs_num", 2)
        if n_convs != 2:
            raise ValueError("libisc_b200's sentiment detector is built for sentiment_convs_num = 2, got %d" % n_convs)
        self.convs = nn.Sequential()
        c = settings["fc_feat_dim"]
        self.feat_dim = c
        for i in range(n_convs):
            self.convs.add_module("conv_%d" % i, nn.Conv2d(c, c // 2, 3, padding=1))
            c //= 2
        self.convs.add_module("dropout", nn.Dropout(settings["dropout_p"]))
        self.convs.add_module("relu", nn.ReLU())
        n = len(sentiment_categories)
        self.senti_conv = nn.Conv2d(c, n, 1)
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.output = nn.Sequential(*[nn.Linear(n, n) for _ in range(settings.get("sentiment_fcs_num", 2))])
        self._packed = None
        self._packed_key = None
        self._ws = None

    def _pack(self, dev):
        lib = _lib.load()
        w0, w1 = self.convs.conv_0.weight, self.convs.conv_1.weight
        key = (dev.index, w0.data_ptr(), w0._version, w1.data_ptr(), w1._version)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        nbytes = lib.isc_senti_packed_bytes(self.feat_dim)
        if nbytes == 0:
            raise RuntimeError("isc_senti_packed_bytes: fc_feat_dim must be a multiple of 256")
        packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        # layout change only: [C_out, C_in, 3, 3] -> [C_out, (ky, kx, C_in)], the K order of the convolution GEMM
        r0 = w0.detach().permute(0, 2, 3, 1).reshape(w0.shape[0], -1).float().contiguous()
        r1 = w1.detach().permute(0, 2, 3, 1).reshape(w1.shape[0], -1).float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.isc_senti_pack(self.feat_dim, _lib.ptr(r0), _lib.ptr(r1), _lib.ptr(packed), nbytes,
                                          _lib.stream_ptr(dev)), "isc_senti_pack")
        self._packed, self._packed_key = packed, key
        return packed

    def _detect(self, features, threshold):
        if self.training:
            raise NotImplementedError("SentimentDetector runs in eval() mode on the B200 path (training it is out of scope)")
        dev = self.senti_conv.weight.device
        if dev.type != "cuda" or not features.is_cuda:
            raise RuntimeError("SentimentDetector only exists as sm_100a CUDA kernels: move the module and the features "
                               "to a B200; there is no CPU fallback")
        lib = _lib.load()
        B = features.shape[0]
        feats = features.reshape(B, 14, 14, self.feat_dim).float().contiguous()
        packed = self._pack(dev)
        n = len(self.sentiment_categories)
        nbytes = lib.isc_senti_workspace_bytes(self.feat_dim, B)
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        output = torch.empty(B, n, **f32)
        maps = torch.empty(B, 14, 14, **f32)
        labels = torch.empty(B, dtype=torch.long, device=dev)
        scores = torch.empty(B, **f32)
        fcs = list(self.output)
        out_w = torch.stack([l.weight.detach().float() for l in fcs]).contiguous() if fcs else None
        out_b = torch.stack([l.bias.detach().float() for l in fcs]).contiguous() if fcs else None
        w1x1 = self.senti_conv.weight.detach().reshape(n, -1).float().contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.isc_senti_detect(
                self.feat_dim, n, _lib.ptr(packed), _lib.ptr(self.convs.conv_0.bias.detach().float().contiguous()),
                _lib.ptr(self.convs.conv_1.bias.detach().float().contiguous()), _lib.ptr(w1x1),
                _lib.ptr(self.senti_conv.bias.detach().float().contiguous()), _lib.ptr(out_w), _lib.ptr(out_b), len(fcs),
                _lib.ptr(feats), B, float(threshold), self.neu_idx, _lib.ptr(output), _lib.ptr(maps), _lib.ptr(labels),
                _lib.ptr(scores), _lib.ptr(self._ws), self._ws.numel(), _lib.stream_ptr(dev)), "isc_senti_detect")
        return output, maps, labels, scores

    def forward(self, features):
        """features [bz, 14, 14, fc_feat_dim] -> (output [bz, n], senti_features [bz, 14, 14])."""
        output, maps, _, _ = self._detect(features, 0.0)
        return output, maps

    def sample(self, features, senti_threshold=0):
        self.eval()
        _, maps, labels, scores = self._detect(features, senti_threshold)
        sentiments = [self.sentiment_categories[i] for i in labels.tolist()]
        return labels, maps, sentiments, scores

    def sample_labels(self, features, senti_threshold=0):
        """``sample(...)[0]`` only: the thresholded labels as a device tensor, without the D2H copy that building the
        list of sentiment NAMES costs (the RL-iteration driver uses nothing else, models/decoder.py:83-84)."""
        self.eval()
        return self._detect(features, senti_threshold)[2]

    def get_optim_criterion(self, lr, weight_decay=0):
        return torch.optim.Adam(self.parameters(), lr=lr, weight_decay=weight_decay), nn.CrossEntropyLoss()
