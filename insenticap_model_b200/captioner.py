"""Drop-in ``Captioner`` whose decode path runs on libisc_b200.so (hand-written sm_100a CUDA).

Mirrors the Python surface of /root/reference/models/captioner.py:121-424 — constructor,
parameter names (``state_dict`` keys load verbatim), ``init_hidden``, ``forward_step``,
``forward(mode=...)``, ``forward_xe`` / ``forward_seq2seq`` / ``forward_rl``, ``sample``,
``get_optim_criterion`` and the post-call attributes — plus a batched ``beam_search``.

PyTorch is plumbing here (parameters, device memory, streams); every FLOP of the path is in the
C-ABI library. There is no CPU and no eager-PyTorch fallback: calling a compute method with the
library missing, on CPU tensors, or on a non-sm_100 device raises.

Training: in ``train()`` mode (or ``eval()`` with gradients enabled, in ``precision="bf16x3"``) forward_xe /
forward_seq2seq / forward_rl(sample_max=0) carry autograd history through the hand-written backward
(isc_train_forward / isc_train_backward): dropout from torch-drawn keep masks, scheduled sampling by Gumbel-max.
Host-tensor inputs to ``beam_search`` / ``sample`` take the pipelined host path and return host tensors that are
complete when the call returns.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


class _ContentAttentionParams(nn.Module):
    """Parameter holder for attention.cont_att.* (reference ContentAttention, captioner.py:12-16)."""

    def __init__(self, s):
        super().__init__()
        self.h2att = nn.Linear(s["rnn_hid_dim"], s["att_hid_dim"])
        self.att_alpha = nn.Linear(s["att_hid_dim"], 1)


class _SentiAttentionParams(nn.Module):
    """Parameter holder for attention.senti_att.* (reference SentiAttention, captioner.py:38-43)."""

    def __init__(self, s):
        super().__init__()
        self.h2word = nn.Linear(s["rnn_hid_dim"], s["att_hid_dim"])
        self.label2word = nn.Linear(s["word_emb_dim"], s["att_hid_dim"])
        self.word_alpha = nn.Linear(s["att_hid_dim"], 1)


class _AttentionParams(nn.Module):
    """Parameter holder for attention.* (reference Attention, captioner.py:65-74)."""

    def __init__(self, s):
        super().__init__()
        self.cont_att = _ContentAttentionParams(s)
        self.senti_att = _SentiAttentionParams(s)
        self.h2att = nn.Linear(s["rnn_hid_dim"], s["att_hid_dim"])
        self.cont2att = nn.Linear(s["feat_emb_dim"], s["att_hid_dim"])
        self.senti2att = nn.Linear(s["feat_emb_dim"], s["att_hid_dim"])
        self.att_alpha = nn.Linear(s["att_hid_dim"], 1)


class XECriterion(nn.Module):
    """Masked NLL (reference XECriterion, captioner.py:427-440)."""

    def forward(self, pred, target, lengths):
        max_len = max(lengths)
        lens = _lib.to_device_async(list(lengths), torch.long, pred.device).unsqueeze(1)
        mask = (torch.arange(max_len, device=pred.device).unsqueeze(0) < lens).to(pred.dtype)
        nll = -pred.gather(2, target.unsqueeze(2)).squeeze(2) * mask
        return nll.sum() / mask.sum()


class _TeacherForced(torch.autograd.Function):
    """Teacher-forced decode with hand-written backward (isc_train_forward / isc_train_backward).

    forward(model, mode, call, *params) -> (logprobs [B,T,V], fc_embedded [B,512], cpt_feats [B,512]); ``params`` are the
    model's parameters in _lib.WEIGHT_FIELDS order (passed so that autograd routes their gradients)."""

    @staticmethod
    def forward(ctx, model, mode, call, *params):
        lib = _lib.load()
        dev = model._device()
        packed = model.pack_weights()
        smp = call.get("sample")  # free-running sampled pass recorded on the tape (isc_train_forward_sample)
        if smp is not None:
            B, n_steps = call["fc"].shape[0], int(smp["n_steps"])
        else:
            B, n_steps = call["inputs"].shape[0], call["inputs"].shape[1] - 1
        d = model._dims(call["n_regions"], call["n_senti"], call.get("att_tile", 1))
        nbytes = lib.isc_train_workspace_bytes(C.byref(d), model._prec, B, n_steps)
        if nbytes == 0:
            _lib.check(-1, "isc_train_workspace_bytes")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)  # holds the tape until backward
        f32 = dict(dtype=torch.float32, device=dev)
        logprobs = torch.empty(B, n_steps, model.vocab_size, **f32)
        fc_emb = torch.zeros(B, 512, **f32)
        cpt = torch.zeros(B, 512, **f32)
        drop = model._dropout_struct(call["dropout"])
        if smp is not None:
            seq = torch.empty(B, n_steps, dtype=torch.long, device=dev)
            seq_lp = torch.empty(B, n_steps, **f32)
            seq_masks = torch.empty(B, n_steps, **f32)
            noise = smp.get("noise")
            with torch.cuda.device(dev):
                _lib.check(lib.isc_train_forward_sample(
                    C.byref(d), _lib.ptr(packed), model._prec, mode, _lib.ptr(call["fc"]), _lib.ptr(call["att"]),
                    _lib.ptr(call["cpt"]), call["cpt"].shape[1] if call["cpt"] is not None else 0, _lib.ptr(call["sw"]),
                    _lib.ptr(call["labels"]), B, n_steps, C.byref(drop) if call["dropout"] else None,
                    1 if noise is not None else 2, _lib.ptr(noise), int(smp.get("seed") or 0), _lib.ptr(seq), _lib.ptr(seq_lp),
                    _lib.ptr(seq_masks), _lib.ptr(logprobs), _lib.ptr(fc_emb), _lib.ptr(cpt) if call["cpt"] is not None else None,
                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "isc_train_forward_sample")
            sos = torch.full((B, 1), model.sos_id, dtype=torch.long, device=dev)
            call["inputs"] = torch.cat([sos, seq], dim=1)  # what the tape was fed: [<SOS>, seq[:, :-1]] (+ one unused column)
            call["gather"] = seq
            ctx.model, ctx.mode, ctx.call, ctx.ws, ctx.dims, ctx.packed = model, mode, call, ws, d, packed
            ctx.save_for_backward(logprobs)
            ctx.param_shapes = [p.shape for p in params]
            ctx.mark_non_differentiable(fc_emb, seq, seq_masks)
            return logprobs.gather(2, seq.unsqueeze(2)).squeeze(2), fc_emb, cpt, seq, seq_masks
        with torch.cuda.device(dev):
            _lib.check(lib.isc_train_forward(
                C.byref(d), _lib.ptr(packed), model._prec, mode, _lib.ptr(call["fc"]), _lib.ptr(call["att"]),
                _lib.ptr(call["cpt"]), call["cpt"].shape[1] if call["cpt"] is not None else 0, _lib.ptr(call["sw"]),
                _lib.ptr(call["labels"]), B, _lib.ptr(call["inputs"]), call["inputs"].shape[1], n_steps,
                C.byref(drop) if call["dropout"] else None, C.byref(call["ss"]) if call.get("ss") is not None else None,
                _lib.ptr(logprobs), _lib.ptr(fc_emb), _lib.ptr(cpt) if call["cpt"] is not None else None, _lib.ptr(ws),
                ws.numel(), _lib.stream_ptr(dev)), "isc_train_forward")
        ctx.model, ctx.mode, ctx.call, ctx.ws, ctx.dims, ctx.packed = model, mode, call, ws, d, packed
        ctx.save_for_backward(logprobs)
        ctx.param_shapes = [p.shape for p in params]
        ctx.mark_non_differentiable(fc_emb)
        if call.get("fused") is not None:
            # fused masked NLL / REINFORCE form: loss = sum coef[b,t] * (-logprobs[b,t,targets[b,t]]); the backward then
            # builds d logits straight from (targets, coef) and never materialises a [B,T,V] gradient
            targets, coef = call["fused"]
            loss = -(logprobs.gather(2, targets.unsqueeze(2)).squeeze(2) * coef).sum()
            return loss, fc_emb, cpt
        if call.get("gather") is not None:
            # REINFORCE form: only the log-probs of the given tokens leave the node, [B,T]; their gradient comes back as
            # [B,T] too and the backward builds d logits from (tokens, -d out) — no dense [B,T,V] gradient (1.6 GB at
            # 2560 rows) is created, zero-filled and scattered into by autograd
            return logprobs.gather(2, call["gather"].unsqueeze(2)).squeeze(2), fc_emb, cpt
        return logprobs, fc_emb, cpt

    @staticmethod
    def backward(ctx, dlogp, _dfc, dcpt, *_unused):
        lib = _lib.load()
        model, call, d = ctx.model, ctx.call, ctx.dims
        dev = model._device()
        (logprobs,) = ctx.saved_tensors
        B, n_steps = call["inputs"].shape[0], call["inputs"].shape[1] - 1
        sizes = [int(torch.Size(sh).numel()) for sh in ctx.param_shapes]
        # A gradient sink (train.FusedClampAdam registers itself as model._grad_sink) owns ONE flat fp32 gradient buffer
        # whose slices are the parameters' .grad: isc_train_backward ACCUMULATES, so it adds straight into those slices and
        # autograd gets None for the parameters — no per-node 88 MB scratch, no AccumulateGrad pass, and the sink knows
        # when slices are final (gradient-ready marks -> bucketed all-reduce under the rest of the backward).
        sink = getattr(model, "_grad_sink", None)
        if sink is not None and not sink.accepts(dev, ctx.param_shapes):
            sink = None
        views, g, off = [], _lib.Grads(), 0
        if sink is not None:
            for field, name in _lib.WEIGHT_FIELDS:
                setattr(g, field, sink.grad_slice(name).data_ptr())
        else:
            flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            for (field, _), n, sh in zip(_lib.WEIGHT_FIELDS, sizes, ctx.param_shapes):
                v = flat[off:off + n].view(sh)
                views.append(v)
                setattr(g, field, v.data_ptr())
                off += n
        no_param_grads = (None,) * len(sizes)
        targets = coef = None
        if call.get("fused") is not None:
            targets, coef = call["fused"]
            coef = (coef * dlogp).contiguous() if dlogp is not None else None  # dlogp is d loss_out here (a scalar)
            if coef is None:
                targets = None
            dlogp = None
        elif call.get("gather") is not None:
            targets = call["gather"] if dlogp is not None else None
            coef = (-dlogp).contiguous() if dlogp is not None else None  # loss form: sum coef * (-logp[target])
            dlogp = None
        else:
            dlogp = dlogp.contiguous() if dlogp is not None else None
        dcpt = dcpt.contiguous() if (dcpt is not None and call["cpt"] is not None) else None
        if dlogp is None and coef is None and dcpt is None:
            if sink is not None:
                sink.node_done(None)
                return (None, None, None) + no_param_grads
            return (None, None, None) + tuple(views)
        drop = model._dropout_struct(call["dropout"])
        marks = sink.node_starting() if sink is not None else None  # events, if this is the iteration's last backward node
        with torch.cuda.device(dev):
            if marks is not None:
                _lib.check(lib.isc_train_backward_marks(marks[0].cuda_event, marks[1].cuda_event), "isc_train_backward_marks")
            _lib.check(lib.isc_train_backward(
                C.byref(d), _lib.ptr(ctx.packed), model._prec, ctx.mode, _lib.ptr(call["fc"]), _lib.ptr(call["att"]),
                _lib.ptr(call["cpt"]), call["cpt"].shape[1] if call["cpt"] is not None else 0, _lib.ptr(call["sw"]),
                _lib.ptr(call["labels"]), B, _lib.ptr(call["inputs"]), call["inputs"].shape[1], n_steps,
                C.byref(drop) if call["dropout"] else None, _lib.ptr(logprobs), _lib.ptr(dlogp), _lib.ptr(targets),
                targets.shape[1] if targets is not None else 0, _lib.ptr(coef), _lib.ptr(dcpt), C.byref(g), _lib.ptr(ctx.ws),
                ctx.ws.numel(), _lib.stream_ptr(dev)), "isc_train_backward")
        ctx.ws = None
        if sink is not None:
            sink.node_done(marks)
            return (None, None, None) + no_param_grads
        return (None, None, None) + tuple(views)


class Captioner(nn.Module):
    def __init__(self, idx2word, sentiment_categories, settings, precision: str = "bf16x3"):
        super().__init__()
        s = settings
        dims = {s["word_emb_dim"], s["feat_emb_dim"], s["rnn_hid_dim"], s["att_hid_dim"]}
        if dims != {512}:
            raise ValueError("libisc_b200 is compiled for word/feat/rnn/att dims of 512, got %s" % sorted(dims))
        if s["fc_feat_dim"] != s["att_feat_dim"]:
            raise ValueError("fc_feat_dim must equal att_feat_dim")
        self.idx2word = idx2word
        self.pad_id = idx2word.index("<PAD>")
        self.unk_id = idx2word.index("<UNK>")
        has_sos = "<SOS>" in idx2word
        self.sos_id = idx2word.index("<SOS>") if has_sos else self.pad_id
        self.eos_id = idx2word.index("<EOS>") if has_sos else self.pad_id
        self.neu_idx = sentiment_categories.index("neutral")
        self.vocab_size = len(idx2word)
        self.settings = dict(s)

        E, F, Hd, A = s["word_emb_dim"], s["feat_emb_dim"], s["rnn_hid_dim"], s["att_hid_dim"]
        self.drop = nn.Dropout(s["dropout_p"])
        self.word_embed = nn.Sequential(nn.Embedding(self.vocab_size, E, padding_idx=self.pad_id), nn.ReLU())
        self.senti_label_embed = nn.Sequential(nn.Embedding(len(sentiment_categories), E), nn.ReLU())
        self.fc_embed = nn.Sequential(nn.Linear(s["fc_feat_dim"], F), nn.ReLU())
        self.cpt2fc = nn.Sequential(nn.Linear(E, F), nn.ReLU())
        self.att_embed = nn.Sequential(nn.Linear(s["att_feat_dim"], F), nn.ReLU())
        self.att_lstm = nn.LSTMCell(Hd + F + E, Hd)
        self.att2att = nn.Sequential(nn.Linear(F, A), nn.ReLU())
        self.senti2att = nn.Sequential(nn.Linear(E, A), nn.ReLU())
        self.attention = _AttentionParams(s)
        self.lang_lstm = nn.LSTMCell(Hd + F, Hd)
        self.classifier = nn.Linear(Hd, self.vocab_size)

        self.n_labels = len(sentiment_categories)
        self.n_regions = 196
        self.num_senti_words = 10
        self.collect_attention_weights = True
        self.fast_features = True  # attention reads fp16 copies of the projected features (tensor-core precisions)
        self.set_precision(precision)
        self._packed = None
        self._packed_key = None
        self._ws = {}
        self.ss_override = None  # tests: {"uniform": [T,B], "noise": [T,B,V]} for scheduled sampling
        self.dropout_override = None  # tests: dict of uint8 keep masks {fc, att, sw, sl, out, scale}
        self.fuse_sampled_tape = True  # forward_rl under autograd: sample on the training tape (one pass) instead of decode + re-score
        self.use_cuda_graph = False  # beam_search: capture the device-side call once and replay it
        # graph replay writes into the SAME output tensors every time; True returns copies (0.4 MB at B = 1024) so that a
        # caller may keep results across calls like with the reference, False hands out the graph's own buffers
        self.graph_outputs_fresh = True
        self._graphs = {}
        self._host_graphs = {}
        self.use_host_graphs = True  # host-tensor beam_search: replay a captured graph per sub-batch (two staging sets)
        self._copy_stream = None
        self.cont_weights = self.senti_weights = self.cont_senti_weights = []
        self.fc_feats = self.cpt_feats = None

    # ------------------------------------------------------------------ plumbing
    def set_precision(self, precision: str):
        if precision not in _lib.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_lib.PRECISIONS))
        self.precision = precision
        self._prec = _lib.PRECISIONS[precision]
        self._packed_key = None

    def _device(self):
        dev = self.classifier.weight.device
        if dev.type != "cuda":
            raise RuntimeError("Captioner parameters are on %s: the decode path only exists as sm_100a CUDA "
                               "kernels, move the module to a B200 (`.cuda()`); there is no CPU fallback" % dev)
        return dev

    def _dims(self, n_regions=None, n_senti=None, att_tile=1):
        return _lib.Dims(self.vocab_size, 512, self.settings["att_feat_dim"],
                         n_regions or self.n_regions, n_senti or (self.num_senti_words + 1), self.n_labels,
                         self.pad_id, self.sos_id, self.eos_id, self.unk_id, int(att_tile))

    def _workspace(self, kind, nbytes, dev):
        key = (kind, dev.index)
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            # a captured CUDA graph holds the raw address of the workspace it was captured with: growing the buffer
            # frees that block, so every graph captured so far is dropped (recaptured on its next use)
            self._graphs.clear()
            self._host_graphs.clear()
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
            self._ws[key] = buf
        return buf

    def pack_weights(self, force=False):
        """Fuse/split the fp32 parameters into the kernel layouts (isc_pack_weights). Re-packs when any
        parameter was modified in place or replaced (optimizer step, load_state_dict)."""
        dev = self._device()
        lib = _lib.load()
        sd = {k: v for k, v in self.named_parameters()}
        key = (self._prec, dev.index) + tuple((p.data_ptr(), p._version) for p in sd.values())
        if not force and self._packed is not None and key == self._packed_key:
            return self._packed
        w = _lib.Weights()
        keep = []
        for field, name in _lib.WEIGHT_FIELDS:
            t = sd[name].detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            setattr(w, field, t.data_ptr())
        d = self._dims()
        nbytes = lib.isc_packed_weights_bytes(C.byref(d), self._prec)
        if nbytes == 0:
            _lib.check(-1, "isc_packed_weights_bytes")
        if self._packed is None or self._packed.numel() < nbytes or self._packed.device != dev:
            self._packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.isc_pack_weights(C.byref(d), C.byref(w), self._prec, _lib.ptr(self._packed), nbytes,
                                            _lib.stream_ptr(dev)), "isc_pack_weights")
        self._packed_key = key
        return self._packed

    def _feat_dtype(self):
        return torch.bfloat16 if self._prec == _lib.PREC_BF16 else torch.float32

    def _make_feats(self, tensors):
        f = _lib.Feats()
        for name in _lib.FEAT_FIELDS:
            t = tensors.get(name)
            setattr(f, name, t.data_ptr() if t is not None else None)
        return f

    def _convert_features(self, x, projected, dtype):
        """Caller-embedded fp32 features -> the representation the kernels read (isc_convert_features)."""
        dev = self._device()
        x = x.float().contiguous()
        out = torch.empty(x.shape, dtype=dtype, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().isc_convert_features(self._prec, 1 if projected else 0, _lib.ptr(x), _lib.ptr(out),
                                                        x.numel(), _lib.stream_ptr(dev)), "isc_convert_features")
        return out

    def _expand_f16(self, x):
        """fp16 device tensor -> fp32 (isc_expand_f16); other dtypes pass through."""
        if x.dtype != torch.float16:
            return x
        dev = self._device()
        x = x.contiguous()
        out = torch.empty(x.shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().isc_expand_f16(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr(dev)), "isc_expand_f16")
        return out

    # ------------------------------------------------------------------ training plumbing
    def _dropout_masks(self, shapes, n_steps, B):
        """uint8 keep masks for the tensors nn.Dropout touches in the reference (captioner.py:182, :200-214,
        :250-262, :296-315). RNG is torch's (plumbing); tests inject masks through ``self.dropout_override``."""
        p = float(self.settings["dropout_p"])
        if self.dropout_override is not None:
            return self.dropout_override
        if not self.training or p <= 0.0:
            return None
        dev = self._device()
        # one Bernoulli(1 - p) kernel per mask, written as uint8 (rand -> compare -> cast was three passes over the
        # [B, regions, 512] mask)
        keep = lambda sh: torch.empty(sh, dtype=torch.uint8, device=dev).bernoulli_(1.0 - p)
        masks = {k: keep(sh) for k, sh in shapes.items()}
        masks["out"] = keep((n_steps, B, 512))
        masks["scale"] = 1.0 / (1.0 - p)
        return masks

    def _dropout_struct(self, masks):
        d = _lib.Dropout()
        if masks:
            for k in ("fc", "att", "sw", "sl", "out"):
                t = masks.get(k)
                setattr(d, k, t.data_ptr() if t is not None else None)
            d.scale = float(masks.get("scale", 1.0))
        else:
            d.scale = 1.0
        return d

    def _params_in_field_order(self):
        sd = dict(self.named_parameters())
        return [sd[name] for _, name in _lib.WEIGHT_FIELDS]

    def _teacher_forced_train(self, mode, fc, att, cpt, sw, labels, inputs, ss_prob, fused=None, gather=None):
        """Differentiable teacher forcing (autograd.Function over the C ABI). ``fused`` = (targets int64 [B,T],
        coef fp32 [B,T]) makes the first result the scalar sum coef * (-logprobs[targets]) instead of the log-probs;
        ``gather`` = tokens int64 [B,T] makes it logprobs[b, t, tokens[b, t]] ([B,T], sparse gradient)."""
        if self._prec != _lib.PREC_BF16X3:
            raise NotImplementedError("the backward pass runs in precision='bf16x3' only")
        dev = self._device()
        B = inputs.shape[0]
        n_steps = inputs.shape[1] - 1
        ss, ss_keep = None, None
        if ss_prob and ss_prob > 0.0 and self.training:  # captioner.py:219: only in train() mode
            o = self.ss_override or {}
            uni = o["uniform"].to(dev).float().contiguous() if "uniform" in o else torch.rand(n_steps, B, device=dev)
            noise = o["noise"].to(dev).float().contiguous() if "noise" in o else None
            ss = _lib.SchedSampling(float(ss_prob), uni.data_ptr(), noise.data_ptr() if noise is not None else None,
                                    int(torch.randint(0, 2 ** 62, (1,)).item()))
            ss_keep = (uni, noise)
        L = att.shape[1] if att is not None else self.n_regions
        S = (sw.shape[1] + 1) if sw is not None else (self.num_senti_words + 1)
        shapes = {"fc": (B, 512), "sl": (B, 512)}
        if att is not None:
            shapes["att"] = (B, L, 512)
        if sw is not None:
            shapes["sw"] = (B, S, 512)
        call = dict(fc=fc, att=att, cpt=cpt, sw=sw, labels=labels, inputs=inputs.long().contiguous(), n_regions=L, n_senti=S,
                    dropout=self._dropout_masks(shapes, n_steps, B), ss=ss, ss_keep=ss_keep, fused=fused,
                    gather=gather.long().clone() if gather is not None else None)
        out, fc_emb, cpt_feats = _TeacherForced.apply(self, mode, call, *self._params_in_field_order())
        self.cont_weights = self.senti_weights = self.cont_senti_weights = []
        return out, fc_emb, (cpt_feats if cpt is not None else None), call

    def prologue(self, fc_feats=None, att_feats=None, cpt_words=None, senti_words=None, senti_labels=None,
                 seq2seq=False, dropout=None):
        """Step-invariant features (isc_prologue). Returns (dict of tensors, B). ``dropout`` is the keep-mask
        dict of ``_dropout_masks`` (training-mode sampling pass) or None."""
        dev = self._device()
        lib = _lib.load()
        packed = self.pack_weights()
        ref = fc_feats if fc_feats is not None else cpt_words
        B = ref.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        t = {}
        n_regions, S = self.n_regions, self.num_senti_words + 1
        bf16_in = False
        if not seq2seq:
            # bf16 features (a bf16 feature shard, dataloader.FeatureShard): precision="bf16" rounds its inputs to bf16
            # anyway, so they go in as they are; every other mode computes on fp32 inputs
            if fc_feats.dtype == torch.float16 or att_feats.dtype == torch.float16:
                # fp16 features (a fp16 feature shard): half the host->device bytes; every fp16 value is an fp32 value and
                # is represented exactly by the split-bf16 operands, so this is the fp32 computation on these inputs
                fc_feats, att_feats = self._expand_f16(fc_feats), self._expand_f16(att_feats)
            bf16_in = (fc_feats.dtype == torch.bfloat16 and att_feats.dtype == torch.bfloat16
                       and self._prec == _lib.PREC_BF16 and not dropout)
            keep = (lambda x: x.contiguous()) if bf16_in else (lambda x: x.float().contiguous())
            fc_feats = keep(fc_feats.reshape(B, -1))
            att_feats = keep(att_feats.reshape(B, -1, att_feats.shape[-1]))
            n_regions = att_feats.shape[1]
            t["att"] = torch.empty(B, n_regions, 512, dtype=self._feat_dtype(), device=dev)
            t["p_att"] = torch.empty_like(t["att"])
            if self.fast_features and self._prec != _lib.PREC_FP32 and not dropout:
                # 16-bit copies for the attention kernel (isc_feats_t::att16 / p_att16 / feat_flags): half the bytes the
                # decode loop streams per step; images outside the fp16 path's exact domain are flagged and read full width
                if self._prec == _lib.PREC_BF16X3:
                    t["att16"] = torch.empty(B, n_regions, 512, dtype=torch.float16, device=dev)
                t["p_att16"] = torch.empty(B, n_regions, 512, dtype=torch.float16, device=dev)
                t["feat_flags"] = torch.empty(B, dtype=torch.int32, device=dev)
        t["fc"] = torch.empty(B, 512, **f32)
        t["pre_gates"] = torch.empty(B, 2048, **f32)
        if cpt_words is not None:
            cpt_words = cpt_words.long().contiguous()
            t["cpt_feats"] = torch.empty(B, 512, **f32)
        if senti_words is not None:
            senti_words = senti_words.reshape(B, -1).long().contiguous()
            S = senti_words.shape[1] + 1
            t["sw"] = torch.empty(B, S, 512, **f32)
            t["p_sw"] = torch.empty(B, S, 512, **f32)
        if senti_labels is not None:
            senti_labels = senti_labels.reshape(B).long().contiguous()
            t["sl"] = torch.empty(B, 512, **f32)
            t["pre_word"] = torch.empty(B, 512, **f32)
        d = self._dims(n_regions, S)
        feats = self._make_feats(t)
        nbytes = lib.isc_prologue_workspace_bytes(C.byref(d), self._prec, B)
        ws = self._workspace("prologue", nbytes, dev)
        if bf16_in:
            with torch.cuda.device(dev):
                _lib.check(lib.isc_prologue_bf16in(
                    C.byref(d), _lib.ptr(packed), self._prec, _lib.ptr(fc_feats), _lib.ptr(att_feats),
                    _lib.ptr(cpt_words), cpt_words.shape[1] if cpt_words is not None else 0,
                    _lib.ptr(senti_words), _lib.ptr(senti_labels), B, C.byref(feats),
                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "isc_prologue_bf16in")
            t["_dims"] = d
            return t, B
        with torch.cuda.device(dev):
            _lib.check(lib.isc_prologue(
                C.byref(d), _lib.ptr(packed), self._prec,
                _lib.ptr(fc_feats) if not seq2seq else None, _lib.ptr(att_feats) if not seq2seq else None,
                _lib.ptr(cpt_words), cpt_words.shape[1] if cpt_words is not None else 0,
                _lib.ptr(senti_words), _lib.ptr(senti_labels), B, 1 if seq2seq else 0, C.byref(feats),
                _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev),
                C.byref(self._dropout_struct(dropout)) if dropout else None), "isc_prologue")
        t["_dims"] = d
        return t, B

    def _decode_ws(self, d, M, dev):
        lib = _lib.load()
        nbytes = lib.isc_decode_workspace_bytes(C.byref(d), self._prec, M)
        return self._workspace("decode", nbytes, dev)

    # ------------------------------------------------------------------ reference API
    def init_hidden(self, bsz):
        w = self.classifier.weight
        return (w.new_zeros([2, bsz, 512]), w.new_zeros([2, bsz, 512]))

    def forward_step(self, it, state, fc_feats, att_feats=None, p_att_feats=None,
                     senti_word_feats=None, p_senti_word_feats=None, senti_labels=None):
        """One decode step on ALREADY-EMBEDDED features (captioner.py:168-186).
        Returns (logprobs [bs, V], (h [2,bs,512], c [2,bs,512]))."""
        dev = self._device()
        lib = _lib.load()
        packed = self.pack_weights()
        M = it.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        t = {"fc": fc_feats.float().contiguous(), "pre_gates": torch.empty(M, 2048, **f32)}
        n_regions, S = self.n_regions, self.num_senti_words + 1
        if att_feats is not None:
            n_regions = att_feats.shape[1]
            t["att"] = self._convert_features(att_feats, False, self._feat_dtype())
            t["p_att"] = self._convert_features(p_att_feats, True, self._feat_dtype())
        if senti_word_feats is not None:
            S = senti_word_feats.shape[1]
            t["sw"] = senti_word_feats.float().contiguous()
            t["p_sw"] = self._convert_features(p_senti_word_feats, True, torch.float32)
        if senti_labels is not None:
            t["sl"] = senti_labels.float().contiguous()
            t["pre_word"] = torch.empty(M, 512, **f32)
        d = self._dims(n_regions, S)
        feats = self._make_feats(t)
        ws = self._decode_ws(d, M, dev)
        h_in, c_in = state[0].float().contiguous(), state[1].float().contiguous()
        h_out, c_out = torch.empty_like(h_in), torch.empty_like(c_in)
        logprobs = torch.empty(M, self.vocab_size, **f32)
        cw = torch.empty(M, n_regions, **f32) if att_feats is not None else None
        sw = torch.empty(M, S, **f32) if senti_word_feats is not None else None
        gw = torch.empty(M, 1, **f32) if (cw is not None and sw is not None) else None
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            _lib.check(lib.isc_hoist(C.byref(d), _lib.ptr(packed), self._prec, M, C.byref(feats), _lib.ptr(ws),
                                     ws.numel(), st), "isc_hoist")
            _lib.check(lib.isc_decode_step(
                C.byref(d), _lib.ptr(packed), self._prec, C.byref(feats), 1, M, _lib.ptr(it.long().contiguous()),
                _lib.ptr(h_in), _lib.ptr(c_in), _lib.ptr(h_out), _lib.ptr(c_out), _lib.ptr(logprobs),
                self.vocab_size, _lib.ptr(cw), _lib.ptr(sw), _lib.ptr(gw), _lib.ptr(ws), ws.numel(), st),
                "isc_decode_step")
        self._step_weights = (cw, sw, gw)
        return logprobs, (h_out, c_out)

    def forward(self, *args, **kwargs):
        mode = kwargs.pop("mode", "xe")
        return getattr(self, "forward_" + mode)(*args, **kwargs)

    def _teacher_forced(self, t, B, inputs):
        dev = self._device()
        lib = _lib.load()
        d = t["_dims"]
        n_steps = inputs.shape[1] - 1
        inputs = inputs.long().contiguous()
        out = torch.empty(B, n_steps, self.vocab_size, dtype=torch.float32, device=dev)
        ws = self._decode_ws(d, B, dev)
        feats = self._make_feats(t)
        with torch.cuda.device(dev):
            _lib.check(lib.isc_teacher_forced(
                C.byref(d), _lib.ptr(self._packed), self._prec, C.byref(feats), B, n_steps, _lib.ptr(inputs),
                inputs.shape[1], _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)),
                "isc_teacher_forced")
        self.cont_weights = self.senti_weights = self.cont_senti_weights = []
        return out

    def _needs_grad(self):
        """True when the call has to build autograd history: train() mode, or eval() with gradients enabled like
        the reference's modules — the latter only in bf16x3, the precision the backward pass exists in
        (other precisions then return plain tensors; use torch.no_grad() for inference either way)."""
        if self.training:
            return True
        return (self._prec == _lib.PREC_BF16X3 and torch.is_grad_enabled()
                and any(p.requires_grad for p in self.parameters()))

    def forward_xe(self, fc_feats, att_feats, cpt_words, captions, senti_labels, ss_prob=0.0):
        """Teacher-forced log-probs [bs, T-1, V] (captioner.py:194-240). With gradients enabled (or in train()
        mode) the result carries autograd history through the hand-written backward (isc_train_backward)."""
        if self._needs_grad():
            B = fc_feats.shape[0]
            fc = fc_feats.reshape(B, -1).float().contiguous()
            att = att_feats.reshape(B, -1, att_feats.shape[-1]).float().contiguous()
            out, self.fc_feats, self.cpt_feats, _ = self._teacher_forced_train(
                _lib.MODE_XE, fc, att, cpt_words.long().contiguous(), None, senti_labels.reshape(B).long().contiguous(),
                captions, ss_prob)
            return out
        t, B = self.prologue(fc_feats, att_feats, cpt_words, None, senti_labels)
        self.fc_feats, self.cpt_feats = t["fc"], t.get("cpt_feats")
        return self._teacher_forced(t, B, captions)

    @staticmethod
    def _nll_coef(captions, lengths, dev):
        """XECriterion's weights (captioner.py:431-440): mask[b,t] = t < lengths[b], normalised by its sum."""
        n_steps = captions.shape[1] - 1
        lens = _lib.to_device_async(list(lengths), torch.long, dev).unsqueeze(1)
        mask = (torch.arange(n_steps, device=dev).unsqueeze(0) < lens).float()
        return captions[:, 1:].long().contiguous(), (mask / mask.sum()).contiguous()

    def xe_loss(self, fc_feats, att_feats, cpt_words, captions, senti_labels, lengths, ss_prob=0.0):
        """``XECriterion()(self(..., mode='xe'), captions[:, 1:], lengths)`` with the loss fused into the backward
        (no [B,T,V] gradient tensor). Sets fc_feats / cpt_feats like forward_xe."""
        B = fc_feats.shape[0]
        targets, coef = self._nll_coef(captions, lengths, self._device())
        loss, self.fc_feats, self.cpt_feats, _ = self._teacher_forced_train(
            _lib.MODE_XE, fc_feats.reshape(B, -1).float().contiguous(),
            att_feats.reshape(B, -1, att_feats.shape[-1]).float().contiguous(), cpt_words.long().contiguous(), None,
            senti_labels.reshape(B).long().contiguous(), captions, ss_prob, fused=(targets, coef))
        return loss

    def seq2seq_loss(self, senti_captions, cpt_words, senti_words, senti_labels, lengths, ss_prob=0.0):
        """Fused-loss form of ``XECriterion()(self(..., mode='seq2seq'), senti_captions[:, 1:], lengths)``."""
        B = senti_captions.shape[0]
        targets, coef = self._nll_coef(senti_captions, lengths, self._device())
        loss, _, _, _ = self._teacher_forced_train(
            _lib.MODE_SEQ2SEQ, None, None, cpt_words.long().contiguous(), senti_words.reshape(B, -1).long().contiguous(),
            senti_labels.reshape(B).long().contiguous(), senti_captions, ss_prob, fused=(targets, coef))
        return loss

    def forward_seq2seq(self, senti_captions, cpt_words, senti_words, senti_labels, ss_prob=0.0):
        """Sentiment-corpus teacher forcing (captioner.py:242-288)."""
        if self._needs_grad():
            B = senti_captions.shape[0]
            out, _, _, _ = self._teacher_forced_train(
                _lib.MODE_SEQ2SEQ, None, None, cpt_words.long().contiguous(), senti_words.reshape(B, -1).long().contiguous(),
                senti_labels.reshape(B).long().contiguous(), senti_captions, ss_prob)
            return out
        t, B = self.prologue(None, None, cpt_words, senti_words, senti_labels, seq2seq=True)
        return self._teacher_forced(t, B, senti_captions)

    def forward_rl(self, fc_feats, att_feats, cpt_words, senti_words, senti_labels, max_seq_len, sample_max,
                   noise=None, seed=None, att_tile=1):
        """Batched greedy (sample_max=1) or sampled (sample_max=0) decode (captioner.py:290-349).

        ``att_tile`` = R > 1 (additive): the batch is TILED — every R consecutive rows of fc_feats / words / labels are the
        same image (SCST with R samples per image) — and ``att_feats`` holds the B / R images once, [B/R, 14, 14, D]. The
        result is what the reference computes on ``att_feats.repeat_interleave(R, 0)``; under autograd the region embedding
        and its weight gradient are then computed once per image instead of once per row (isc_dims_t::att_tile).

        Sampling is Gumbel-max: argmax(logprobs + g). ``noise`` [T,B,V] supplies g explicitly (parity
        tests); otherwise g comes from a counter-based generator keyed by ``seed`` (drawn from torch's
        global generator when None, so torch.manual_seed makes it reproducible)."""
        dev = self._device()
        lib = _lib.load()
        with_grad = self._needs_grad() and torch.is_grad_enabled() and not sample_max
        R = int(att_tile)
        if R > 1 and not (with_grad and not self.collect_attention_weights and self.fuse_sampled_tape):
            att_feats, R = att_feats.repeat_interleave(R, dim=0), 1  # every other path takes the tiled tensor itself
        masks = None
        if with_grad or (self.training and float(self.settings["dropout_p"]) > 0.0) or self.dropout_override is not None:
            B0, L0 = fc_feats.shape[0], att_feats.reshape(att_feats.shape[0], -1, att_feats.shape[-1]).shape[1]
            masks = self._dropout_masks({"fc": (B0, 512), "sl": (B0, 512), "att": (B0, L0, 512),
                                         "sw": (B0, senti_words.reshape(B0, -1).shape[1] + 1, 512)}, int(max_seq_len), B0)
        if with_grad and not self.collect_attention_weights and self.fuse_sampled_tape:
            # REINFORCE's sampled pass (models/decoder.py:86-88 runs it under autograd) in ONE pass: the decoder runs free on
            # the training tape (isc_train_forward_sample), so the tokens, their log-probs and the activations the backward
            # needs come out together — no sampling decode followed by a teacher-forced re-scoring of its tokens
            B0 = fc_feats.shape[0]
            T = int(max_seq_len)
            if noise is not None:
                noise = noise.to(dev).float().contiguous()
                assert noise.shape == (T, B0, self.vocab_size)
            elif seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            assert B0 % R == 0 and att_feats.shape[0] * R == B0, "att_tile: att_feats must hold B / att_tile images"
            att3 = att_feats.reshape(B0 // R, -1, att_feats.shape[-1]).float().contiguous()
            sw2 = senti_words.reshape(B0, -1).long().contiguous()
            call = dict(fc=fc_feats.reshape(B0, -1).float().contiguous(), att=att3, cpt=cpt_words.long().contiguous(), sw=sw2,
                        labels=senti_labels.reshape(B0).long().contiguous(), n_regions=att3.shape[1], n_senti=sw2.shape[1] + 1,
                        dropout=masks, ss=None, fused=None, sample=dict(n_steps=T, noise=noise, seed=seed), att_tile=R)
            lps, self.fc_feats, self.cpt_feats, seq, seq_masks = _TeacherForced.apply(self, _lib.MODE_RL, call,
                                                                                      *self._params_in_field_order())
            self.cont_weights = self.senti_weights = self.cont_senti_weights = []
            executed = seq_masks.sum(0, keepdim=True).gt(0).to(lps.dtype)  # steps after the whole-batch stop never ran
            return seq, lps * executed, seq_masks
        t, B = self.prologue(fc_feats, att_feats, cpt_words, senti_words, senti_labels, dropout=masks)
        self.fc_feats, self.cpt_feats = t["fc"], t.get("cpt_feats")
        d = t["_dims"]
        T = int(max_seq_len)
        f32 = dict(dtype=torch.float32, device=dev)
        seq = torch.empty(B, T, dtype=torch.long, device=dev)
        lps = torch.empty(B, T, **f32)
        seq_masks = torch.empty(B, T, **f32)
        L, S = d.n_regions, d.n_senti
        cw = sw = gw = None
        if self.collect_attention_weights:
            cw = torch.zeros(B, T, L, **f32)
            sw = torch.zeros(B, T, S, **f32)
            gw = torch.zeros(B, T, **f32)
        if sample_max:
            mode = 0
        elif noise is not None:
            mode = 1
            noise = noise.to(dev).float().contiguous()
            assert noise.shape == (T, B, self.vocab_size)
        else:
            mode = 2
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        ws = self._decode_ws(d, B, dev)
        feats = self._make_feats(t)
        with torch.cuda.device(dev):
            _lib.check(lib.isc_decode_greedy(
                C.byref(d), _lib.ptr(self._packed), self._prec, C.byref(feats), B, T, mode, _lib.ptr(noise),
                int(seed or 0), _lib.ptr(seq), _lib.ptr(lps), _lib.ptr(seq_masks), _lib.ptr(cw), _lib.ptr(sw),
                _lib.ptr(gw), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev),
                _lib.ptr(masks["out"]) if masks and masks.get("out") is not None else None,
                float(masks["scale"]) if masks else 1.0), "isc_decode_greedy")
        masks_t = seq_masks
        if self.collect_attention_weights:
            # the reference's lists hold one entry per EXECUTED step (it breaks when every row is done)
            steps = int(seq_masks.sum(0).gt(0).sum().item())
            self.cont_weights = cw[:, :steps].reshape(B, steps * L)
            self.senti_weights = sw[:, :steps].reshape(B, steps * S)
            self.cont_senti_weights = gw[:, :steps]
        else:
            self.cont_weights = self.senti_weights = self.cont_senti_weights = []
        if with_grad:
            # REINFORCE (models/decoder.py:86-88 runs this pass under autograd): re-score the sampled tokens teacher-forced
            # with the same dropout masks — the same computation, now with the tape the backward needs
            sos = torch.full((B, 1), self.sos_id, dtype=torch.long, device=dev)
            inputs = torch.cat([sos, seq], dim=1)  # feeds [SOS, seq[:, :-1]]
            saved, self.dropout_override = self.dropout_override, masks
            try:
                out, self.fc_feats, self.cpt_feats, _ = self._teacher_forced_train(
                    _lib.MODE_RL, fc_feats.reshape(B, -1).float().contiguous(),
                    att_feats.reshape(B, -1, att_feats.shape[-1]).float().contiguous(), cpt_words.long().contiguous(),
                    senti_words.reshape(B, -1).long().contiguous(), senti_labels.reshape(B).long().contiguous(), inputs, 0.0,
                    gather=seq)
            finally:
                self.dropout_override = saved
            executed = seq_masks.sum(0, keepdim=True).gt(0).to(out.dtype)  # steps after the whole-batch stop never ran
            lps = out * executed
        return seq, lps, masks_t

    def beam_search(self, fc_feats, att_feats, senti_words=None, senti_labels=None, beam_size=3,
                    decoding_constraint=1, max_seq_len=16, host_chunk=256):
        """Batched beam search: Captioner.sample (captioner.py:351-420) for B images at once.
        Returns tokens int64 [B,K,T] (EOS included, 0 padded), scores float64 [B,K], lengths int32 [B,K].

        Inputs may be CUDA tensors, or HOST tensors (pinned for speed): host batches are cut into sub-batches of
        ``host_chunk`` images whose H2D copies run on a copy stream while the previous sub-batch decodes
        (images are independent, so the result is bit-identical to the one-batch call); the results then come
        back as pinned host tensors. With ``self.use_cuda_graph`` the whole device-side call (~190 launches) is
        captured once per (input buffers, shapes) and replayed; outputs are then reused by the next replay."""
        if not fc_feats.is_cuda:
            return self._beam_search_host(fc_feats, att_feats, senti_words, senti_labels, beam_size,
                                          decoding_constraint, max_seq_len, host_chunk)
        if not self.use_cuda_graph:
            return self._beam_search_device(fc_feats, att_feats, senti_words, senti_labels, beam_size,
                                            decoding_constraint, max_seq_len)
        args = (fc_feats, att_feats, senti_words, senti_labels)
        self.pack_weights()
        key = tuple((a.data_ptr(), tuple(a.shape), a.dtype) if a is not None else None for a in args) + (
            int(beam_size), int(bool(decoding_constraint)), int(max_seq_len), self._packed_key)
        hit = self._graphs.get(key)
        if hit is None:
            if len(self._graphs) >= 8:
                self._graphs.clear()
            # eager run first: sizes the workspaces outside the capture and surfaces argument errors
            self._beam_search_device(*args, beam_size, decoding_constraint, max_seq_len)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._beam_search_device(*args, beam_size, decoding_constraint, max_seq_len)
            # args, workspaces and packed weights kept alive: the graph holds their addresses
            hit = self._graphs[key] = (g, out, args, tuple(self._ws.values()), self._packed)
        hit[0].replay()
        if self.graph_outputs_fresh:
            return tuple(o.clone() for o in hit[1])
        return hit[1]

    def _host_graph(self, inputs, n, K, constraint, T, dev):
        """Per (sub-batch shape, dtypes, beam, length) TWO staging sets — device input tensors, device outputs and a CUDA
        graph of the whole device-side call on them (~190 launches): the pipelined host path replays a graph per sub-batch
        instead of launching eagerly (no per-launch host cost, and the copy engine only ever waits for a free set)."""
        self.pack_weights()
        key = tuple((tuple(x.shape[1:]), x.dtype) if x is not None else None for x in inputs) + (
            int(n), K, int(bool(constraint)), T, self._packed_key)
        hit = self._host_graphs.get(key)
        if hit is None:
            if len(self._host_graphs) >= 8:
                self._host_graphs.clear()
            sets = []
            for _ in range(2):
                stage = [torch.empty((n,) + tuple(x.shape[1:]), dtype=x.dtype, device=dev) if x is not None else None
                         for x in inputs]
                for t in stage:
                    if t is not None:
                        t.zero_()  # valid word ids / finite features for the sizing and capture runs
                if stage[3] is not None:
                    stage[3].fill_(0)
                out = (torch.empty(n, K, T, dtype=torch.long, device=dev), torch.empty(n, K, dtype=torch.float64, device=dev),
                       torch.empty(n, K, dtype=torch.int32, device=dev))
                self._beam_search_device(*stage, K, constraint, T, out=out)  # sizes the workspaces outside the capture
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._beam_search_device(*stage, K, constraint, T, out=out)
                sets.append((stage, out, g))
            hit = self._host_graphs[key] = (sets, tuple(self._ws.values()), self._packed)
        return hit[0]

    def _beam_search_host(self, fc_feats, att_feats, senti_words, senti_labels, beam_size, decoding_constraint,
                          max_seq_len, host_chunk):
        dev = self._device()
        B, K, T = fc_feats.shape[0], int(beam_size), int(max_seq_len)
        n = min(int(host_chunk), B)
        inputs = (fc_feats, att_feats, senti_words, senti_labels)
        cur = torch.cuda.current_stream(dev)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        tokens = torch.empty(B, K, T, dtype=torch.long, device=dev)
        scores = torch.empty(B, K, dtype=torch.float64, device=dev)
        lengths = torch.empty(B, K, dtype=torch.int32, device=dev)
        chunks = [(lo, min(B, lo + n)) for lo in range(0, B, n)]
        # (Measured and dropped: halving the LAST sub-batch, whose decode no copy hides — one more graph shape, and the
        # two half-size decodes take longer than the tail they were meant to shorten: 55.8 k -> 50.8 k captions/s.)
        sets = None
        if self.use_host_graphs:
            # largest shape first: sizing a larger one later would reallocate the workspaces under the graphs captured before
            sizes = sorted({n} | {hi - lo for lo, hi in chunks if hi - lo in (n, n // 2)}, reverse=True)
            sets = {sz: self._host_graph(inputs, sz, K, decoding_constraint, T, dev) for sz in sizes}
        cs.wait_stream(cur)
        if sets is None:
            staged = []
            for lo, hi in chunks:  # enqueue every H2D copy up front: the copy engine runs ahead
                with torch.cuda.stream(cs):
                    d = [x[lo:hi].to(dev, non_blocking=True) if x is not None else None for x in inputs]
                    ev = torch.cuda.Event()
                    ev.record(cs)
                staged.append((lo, hi, d, ev))
            for lo, hi, d, ev in staged:
                cur.wait_event(ev)
                for x in d:
                    if x is not None:
                        x.record_stream(cur)
                self._beam_search_device(*d, K, decoding_constraint, T, out=(tokens[lo:hi], scores[lo:hi], lengths[lo:hi]))
        else:
            # two staging sets: the copy of sub-batch i + 2 waits (on the device) until the replay of sub-batch i has read
            # its set; with the decode of a sub-batch shorter than its copy, the copy engine never idles
            ready, consumed, which = {}, {}, {}
            seen = {}
            for i, (lo, hi) in enumerate(chunks):  # staging set of sub-batch i: (graph shape, alternating 0 / 1 per shape)
                sz = n // 2 if (hi - lo <= n // 2 and n // 2 in sets) else n
                which[i] = (sz, seen.get(sz, 0) % 2)
                seen[sz] = seen.get(sz, 0) + 1

            def copy_in(i):
                lo, hi = chunks[i]
                sz, par = which[i]
                stage = sets[sz][par][0]
                with torch.cuda.stream(cs):
                    if which[i] in consumed:
                        cs.wait_event(consumed[which[i]])
                    for dst, src in zip(stage, inputs):
                        if dst is not None:
                            dst[:hi - lo].copy_(src[lo:hi], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                ready[i] = ev

            issued = 0
            def top_up(upto):  # copies are issued in order, at most two sub-batches ahead of the replays
                nonlocal issued
                while issued < len(chunks) and issued <= upto:
                    copy_in(issued)
                    issued += 1

            top_up(1)
            for i, (lo, hi) in enumerate(chunks):
                sz, par = which[i]
                stage, out, g = sets[sz][par]
                cur.wait_event(ready[i])
                g.replay()  # a short last sub-batch decodes the set's stale tail rows too; only [lo, hi) is kept
                tokens[lo:hi].copy_(out[0][:hi - lo])
                scores[lo:hi].copy_(out[1][:hi - lo])
                lengths[lo:hi].copy_(out[2][:hi - lo])
                ev = torch.cuda.Event()
                ev.record(cur)
                consumed[which[i]] = ev
                top_up(i + 2)
        outs = []
        for x in (tokens, scores, lengths):
            h = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
            h.copy_(x, non_blocking=True)
            outs.append(h)
        # host tensors are handed to code that reads them with the CPU: the call returns only when the D2H copies
        # have landed (an event on the current stream, not a device-wide synchronize)
        done = torch.cuda.Event()
        done.record(cur)
        done.synchronize()
        return tuple(outs)

    def _beam_search_device(self, fc_feats, att_feats, senti_words, senti_labels, beam_size, decoding_constraint,
                            max_seq_len, out=None):
        dev = self._device()
        lib = _lib.load()
        t, B = self.prologue(fc_feats, att_feats, None, senti_words, senti_labels)
        d = t["_dims"]
        K, T = int(beam_size), int(max_seq_len)
        if out is None:
            tokens = torch.empty(B, K, T, dtype=torch.long, device=dev)
            scores = torch.empty(B, K, dtype=torch.float64, device=dev)
            lengths = torch.empty(B, K, dtype=torch.int32, device=dev)
        else:
            tokens, scores, lengths = out
        ws = self._decode_ws(d, B * K, dev)
        feats = self._make_feats(t)
        with torch.cuda.device(dev):
            _lib.check(lib.isc_decode_beam(
                C.byref(d), _lib.ptr(self._packed), self._prec, C.byref(feats), B, K, T,
                1 if decoding_constraint else 0, _lib.ptr(tokens), _lib.ptr(scores), _lib.ptr(lengths),
                _lib.ptr(ws), ws.numel(), _lib.stream_ptr(dev)), "isc_decode_beam")
        return tokens, scores, lengths

    def sample(self, fc_feat, att_feat, senti_words=None, senti_label=None, beam_size=3,
               decoding_constraint=1, max_seq_len=16):
        """Single-image beam search with the reference's signature and return types
        (captioner.py:351-420): (list[str] of K captions, list[float] of K scores)."""
        self.eval()
        fc = fc_feat.reshape(1, -1)
        att = att_feat.reshape(1, -1, att_feat.shape[-1])
        sw = senti_words.reshape(1, -1) if senti_words is not None else None
        sl = senti_label.reshape(1) if senti_words is not None else None
        tokens, scores, lengths = self.beam_search(fc, att, sw, sl, beam_size, decoding_constraint, max_seq_len)
        tokens, scores, lengths = tokens[0].tolist(), scores[0].tolist(), lengths[0].tolist()
        captions = [" ".join(self.idx2word[w] for w in tokens[k][:lengths[k]] if w != self.eos_id)
                    for k in range(len(tokens))]
        self.cont_weights = self.senti_weights = self.cont_senti_weights = []
        return captions, scores

    def get_optim_criterion(self, lr, weight_decay=0):
        return (torch.optim.Adam(self.parameters(), lr=lr, weight_decay=weight_decay),
                XECriterion(), nn.MSELoss())
