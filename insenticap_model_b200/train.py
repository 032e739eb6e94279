"""Data-parallel training of the captioner on the B200 path: the inner loops of the reference's
train_xe.py:144-196 (XE epoch body) and models/decoder.py:52-176 (`Detector.forward`, the SCST iteration), with
  * forward/backward through libisc_b200.so (Captioner.forward_xe / forward_seq2seq / forward_rl under autograd),
  * the CIDEr-D reward computed on the device (reward.py) — sampled ids, greedy ids, rewards and the REINFORCE
    loss never leave HBM (the reference round-trips through numpy, self_critical/utils.py:59-60, decoder.py:103),
  * ONE collective per iteration: a summing all-reduce of the flat fp32 gradient (22.06 M elements at V = 10 000)
    over torch.distributed (NCCL over NVLink on a B200 box, gloo in the CPU tests), then the reference's element-wise
    clamp to +-grad_clip (train_xe.py:19-23) and Adam fused in one kernel over the flat buffers (isc_adam_step).
The reference is single-device: data parallelism is new; images shard by rank with
no other communication. Losses are per-batch means, so ranks must hold equal batch sizes for the average of
gradients to equal the single-large-batch gradient.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .captioner import XECriterion
from .reward import RewardCriterion, get_self_critical_reward


def flatten_parameters(module: nn.Module):
    """Re-home every parameter (and a matching .grad) as a view into one flat fp32 buffer each, in
    named_parameters() order. Returns (flat_params, flat_grads). Works on any device."""
    params = [p for _, p in module.named_parameters()]
    n = sum(p.numel() for p in params)
    dev = params[0].device
    flat_p = torch.empty(n, dtype=torch.float32, device=dev)
    flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        k = p.numel()
        flat_p[off:off + k].copy_(p.data.reshape(-1))
        p.data = flat_p[off:off + k].view_as(p)
        p.grad = flat_g[off:off + k].view_as(p)
        off += k
    return flat_p, flat_g


def allreduce_gradients(flat_grads: torch.Tensor, group=None) -> int:
    """The path's one exchange step: sum the flat gradient over the ranks. Returns the world size (1 without an
    initialised process group); the division by it is folded into the optimizer kernel (grad_scale)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return world


class _HyperView(dict):
    """The single param group of FusedClampAdam: a dict whose 'lr' / 'weight_decay' / 'betas' / 'eps' entries write
    through to the optimizer (``for g in optim.param_groups: g['lr'] = lr`` as in train_xe.py:132-133)."""
    _KEYS = ("lr", "weight_decay", "betas", "eps")

    def __init__(self, opt):
        super().__init__()
        self._opt = opt
        for k in self._KEYS:
            dict.__setitem__(self, k, getattr(opt, k))
        dict.__setitem__(self, "params", list(opt.model.parameters()))

    def __getitem__(self, k):
        return getattr(self._opt, k) if k in self._KEYS else dict.__getitem__(self, k)

    def __setitem__(self, k, v):
        if k in self._KEYS:
            setattr(self._opt, k, v)
        dict.__setitem__(self, k, v)


class FusedClampAdam:
    """clip_gradient(optimizer, grad_clip) + torch.optim.Adam.step() (train_xe.py:19-23, :191-192;
    models/decoder.py:168-170) as one kernel over the flat parameter / gradient / moment buffers."""

    def __init__(self, model, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, grad_clip=0.1, group=None):
        self.model = model
        self.lr, self.weight_decay, self.betas, self.eps, self.grad_clip = lr, weight_decay, betas, eps, grad_clip
        self.group = group
        self.flat_p, self.flat_g = flatten_parameters(model)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.steps = 0
        model._packed_key = None
        # gradient sink (captioner._TeacherForced.backward): name -> slice of the flat gradient, bucket boundaries
        self._slices, off = {}, 0
        for name, p in model.named_parameters():
            self._slices[name] = (off, p.numel(), tuple(p.shape))
            off += p.numel()
        self._pending_nodes = 0   # captioner backward nodes still to run in this iteration (begin_iteration)
        self._async = []          # (work handle) of all-reduces already in flight
        self._reduced = []        # [(lo, hi)] ranges of flat_g already handed to the collective
        self._comm_stream = None
        self._marks = None
        self.overlap = True
        self.overlap_note = None
        if hasattr(model, "forward_rl"):  # a Captioner: its backward adds straight into flat_g
            model._grad_sink = self

    # ---- gradient sink -------------------------------------------------------------------------------------------
    def accepts(self, dev, shapes):
        return self.flat_g.device == dev and self.flat_g.is_cuda

    def grad_slice(self, name):
        off, n, _ = self._slices[name]
        return self.flat_g[off:off + n]

    def begin_iteration(self, n_nodes):
        """Called by the iteration drivers before backward(): ``n_nodes`` captioner forward passes carry gradients. The
        LAST of their backward nodes to run gets the gradient-ready marks (isc_train_backward_marks)."""
        self._pending_nodes = int(n_nodes)
        self._async, self._reduced = [], []

    def _world(self):
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def node_starting(self):
        self._pending_nodes -= 1
        if self._pending_nodes != 0 or not self.overlap or self._world() <= 1:
            return None
        if self._marks is None:
            self._marks = (torch.cuda.Event(), torch.cuda.Event())
            for ev in self._marks:
                ev.record()  # creates the CUDA events (torch makes them lazily); the library re-records them
            self._comm_stream = torch.cuda.Stream(self.flat_g.device)
        return self._marks

    def _bucket_bounds(self):
        """Flat-buffer ranges that are final at mark 0 / mark 1 of isc_train_backward (named_parameters order: ...,
        att_lstm.*, att2att, senti2att, attention.cont_att.*, attention.senti_att.*, attention.{h2att,...}, lang_lstm.*,
        classifier.*): the tail from attention.h2att on, then the att_lstm block."""
        tail_lo = self._slices["attention.h2att.weight"][0]
        a_lo = self._slices["att_lstm.weight_ih"][0]
        last = self._slices["att_lstm.bias_hh"]
        return (tail_lo, self.flat_g.numel()), (a_lo, last[0] + last[1])

    def node_done(self, marks):
        """After the node's launches are queued: all-reduce each final bucket on the side stream as soon as its event
        fires, i.e. under the remaining backward work of this node."""
        if marks is None:
            return
        for ev, (lo, hi) in zip(marks, self._bucket_bounds()):
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                self._async.append(dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._reduced.append((lo, hi))

    # -- torch.optim.Optimizer surface the reference's scripts use: param_groups[...]['lr'] for the learning-rate decay
    #    (train_xe.py:130-133), state_dict()/load_state_dict() for checkpoint save / resume (train_xe.py:53, :245)
    @property
    def param_groups(self):
        return [self._group]

    @property
    def _group(self):
        g = self.__dict__.get("_group_dict")
        if g is None:
            g = self.__dict__["_group_dict"] = _HyperView(self)
        return g

    def state_dict(self):
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "steps": int(self.steps),
                "lr": float(self.lr), "weight_decay": float(self.weight_decay), "betas": tuple(self.betas),
                "eps": float(self.eps), "grad_clip": float(self.grad_clip)}

    def load_state_dict(self, sd):
        for k in ("exp_avg", "exp_avg_sq"):
            t = sd[k]
            if t.numel() != self.flat_p.numel():
                raise ValueError("optimizer state %s has %d elements, the model has %d" % (k, t.numel(), self.flat_p.numel()))
            getattr(self, k).copy_(t.to(self.flat_p.device).reshape(-1))
        self.steps = int(sd["steps"])
        self.lr, self.weight_decay = float(sd["lr"]), float(sd["weight_decay"])
        self.betas, self.eps, self.grad_clip = tuple(sd["betas"]), float(sd["eps"]), float(sd["grad_clip"])

    def zero_grad(self):
        self.flat_g.zero_()

    def _gather_grads(self):
        """autograd accumulates into p.grad in place when it exists, so the views stay attached; parameters
        that got no gradient keep their zeros."""
        off = 0
        for p in self.model.parameters():
            k = p.numel()
            if p.grad is None:
                p.grad = self.flat_g[off:off + k].view_as(p)
            elif p.grad.data_ptr() != self.flat_g[off:off + k].data_ptr():
                self.flat_g[off:off + k].copy_(p.grad.reshape(-1))
                p.grad = self.flat_g[off:off + k].view_as(p)
            off += k

    def step(self):
        if not self.flat_p.is_cuda:
            raise RuntimeError("FusedClampAdam.step runs on the GPU only (isc_adam_step); there is no CPU fallback")
        self._gather_grads()
        if self._reduced and self._pending_nodes != 0:
            raise RuntimeError("FusedClampAdam: begin_iteration() announced fewer backward nodes than ran; gradients were "
                               "added after their bucket's all-reduce had started")
        if self._reduced:
            # buckets already in flight (started inside the last backward node): reduce what is left — the head of the
            # buffer and the slice between the two buckets —, then join the asynchronous ones
            world = self._world()
            done = sorted(self._reduced)
            pos, rest = 0, []
            for lo, hi in done:
                if lo > pos:
                    rest.append((pos, lo))
                pos = hi
            if pos < self.flat_g.numel():
                rest.append((pos, self.flat_g.numel()))
            for lo, hi in rest:
                dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
            for work in self._async:
                work.wait()
            n_async = sum(hi - lo for lo, hi in done)
            self.overlap_note = "all-reduce in %d buckets: %.0f %% of the gradient started inside the backward" % (
                len(done) + len(rest), 100.0 * n_async / self.flat_g.numel())
            self._async, self._reduced = [], []
        else:
            world = allreduce_gradients(self.flat_g, self.group)
        self.steps += 1
        lib = _lib.load()
        with torch.cuda.device(self.flat_p.device):
            _lib.check(lib.isc_adam_step(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.exp_avg),
                                         _lib.ptr(self.exp_avg_sq), self.flat_p.numel(), float(self.grad_clip), float(self.lr),
                                         float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                                         self.steps, 1.0 / world, _lib.stream_ptr(self.flat_p.device)), "isc_adam_step")
        self.model._packed_key = None  # parameters changed behind torch's version counters: repack on next use


def xe_iteration(model, optim, batch, seq2seq_batch=None, ss_prob=0.0, fused_loss=True):
    """One iteration of train_xe.py:150-192 on device tensors.
    batch = (fc_feats, att_feats, captions, lengths, cpt_words, senti_labels); seq2seq_batch =
    (captions, lengths, cpt_words, senti_words, senti_labels) or None. Returns the loss values (device scalars).
    ``fused_loss``: Captioner.xe_loss / seq2seq_loss (masked NLL folded into the backward) instead of materialising
    the [B,T,V] log-probs' gradient; same value, same gradients."""
    xe_crit, da_crit = XECriterion(), nn.MSELoss()
    fc, att, caps, lengths, cpts, labels = batch
    optim.zero_grad()
    if hasattr(optim, "begin_iteration"):
        optim.begin_iteration(1 + (seq2seq_batch is not None))
    if fused_loss:
        xe_loss = model.xe_loss(fc, att, cpts, caps, labels, lengths, ss_prob)
    else:
        xe_loss = xe_crit(model(fc, att, cpts, caps, labels, ss_prob, mode="xe"), caps[:, 1:], lengths)
    da_loss = da_crit(model.cpt_feats, model.fc_feats.detach())
    all_loss = xe_loss + da_loss
    out = {"xe_loss": xe_loss.detach(), "da_loss": da_loss.detach()}
    if seq2seq_batch is not None:
        s_caps, s_lengths, s_cpts, s_sentis, s_labels = seq2seq_batch
        if fused_loss:
            s2s = model.seq2seq_loss(s_caps, s_cpts, s_sentis, s_labels, s_lengths, ss_prob)
        else:
            s2s = xe_crit(model(s_caps, s_cpts, s_sentis, s_labels, ss_prob, mode="seq2seq"), s_caps[:, 1:], s_lengths)
        all_loss = all_loss + s2s
        out["seq2seq_loss"] = s2s.detach()
    all_loss.backward()
    optim.step()
    out["all_loss"] = all_loss.detach()
    return out


def rl_iteration(model, optim, scorer, batch, max_seq_len=16, seq2seq_batch=None, samples_per_image=1, xe_ss_prob=0.0,
                 seq_flag=1.0, extra_reward=None):
    """One 'fact' iteration of Detector.forward (models/decoder.py:62-170) with the sentiment labels given (the
    detector that predicts them is out of scope) and the classifier reward supplied by
    ``extra_reward(sample, greedy) -> [B,T] tensor`` or omitted.
    batch = (fns, fc_feats, att_feats, captions, lengths, cpt_words, senti_words, senti_labels, ground_truth).
    ``samples_per_image`` tiles the batch (BASELINE cfg4: 5 samples per image; the reference draws 1)."""
    fns, fc, att, caps, lengths, cpts, sentis, labels, ground_truth = batch
    rl_crit, xe_crit, da_crit = RewardCriterion(), XECriterion(), nn.MSELoss()
    optim.zero_grad()
    if hasattr(optim, "begin_iteration"):
        optim.begin_iteration(1 + (caps is not None) + (seq2seq_batch is not None))
    R = int(samples_per_image)
    rep = (lambda x: x.repeat_interleave(R, dim=0)) if R > 1 else (lambda x: x)
    s_fns = [fn for fn in fns for _ in range(R)]
    # nobody reads the per-step attention weights here (Detector.forward does the same): without them the sampled pass runs
    # once, on the training tape
    collect, model.collect_attention_weights = model.collect_attention_weights, False
    try:
        # the tiles share their image's region features: att goes in once per image (att_tile), every other input tiled
        sample, sample_lp, seq_masks = model(rep(fc), att, rep(cpts), rep(sentis), rep(labels), max_seq_len, 0, mode="rl",
                                             att_tile=R)
    finally:
        model.collect_attention_weights = collect
    da_loss = da_crit(model.cpt_feats, model.fc_feats.detach())
    was_training = model.training
    model.eval()
    with torch.no_grad():
        greedy, _, _ = model(fc, att, cpts, sentis, labels, max_seq_len, 1, mode="rl")
    model.train(was_training)
    rewards = get_self_critical_reward(sample, rep(greedy), s_fns, ground_truth, model.sos_id, model.eos_id, scorer,
                                       as_tensor=True).float()
    if extra_reward is not None:
        rewards = rewards + extra_reward(sample, rep(greedy))
    cap_loss = rl_crit(sample_lp, seq_masks, rewards)
    out = {"fact_reward": rewards[:, 0].mean().detach(), "cap_loss": cap_loss.detach(), "da_loss": da_loss.detach()}
    total = cap_loss + da_loss
    if caps is not None:
        pred = model(fc, att, cpts, caps, labels, xe_ss_prob, mode="xe")
        xe_loss = xe_crit(pred, caps[:, 1:], lengths)
        total = total + xe_loss
        out["xe_loss"] = xe_loss.detach()
    if seq2seq_batch is not None:
        s_caps, s_lengths, s_cpts, s_sentis, s_labels = seq2seq_batch
        pred2 = model(s_caps, s_cpts, s_sentis, s_labels, 0.0, mode="seq2seq")
        s2s = seq_flag * xe_crit(pred2, s_caps[:, 1:], s_lengths)
        total = total + s2s
        out["seq2seq_loss"] = s2s.detach()
    total.backward()
    optim.step()
    out["all_loss"] = total.detach()
    return out
