"""Drop-in ``Detector`` (the RL-iteration driver of /root/reference/models/decoder.py:21-192) over the B200 captioner.

Same constructor, attributes, ``set_*`` methods, ``forward(data, data_type, training)`` and ``sample(...)``. The
captioner passes (sampled decode under autograd, greedy baseline, XE and seq2seq teacher forcing), the CIDEr-D
reward and the clamp + Adam step run on libisc_b200.so; sampled ids, greedy ids, rewards and losses stay on the
device (the reference hops through numpy at self_critical/utils.py:59-60 and models/decoder.py:103).

The image sentiment detector and the sentence sentiment classifier are passed in: ``sentiment_detector.SentimentDetector``
and ``sent_senti_cls.SentenceSentimentClassifier`` keep them on libisc_b200.so too,
and any torch modules with the reference's interfaces (``senti_detector.sample(att_feats, threshold) -> (labels, ...)``,
``sent_senti_cls(seqs, lengths) -> (pred [B,3], att_weights [B,max_len])``) — the reference's own classes included — work
as well. Without them, ``data_type='senti'`` batches (which carry their labels) still run, with the classifier reward
switched off.
"""
from __future__ import annotations

from collections import defaultdict

import torch
import torch.nn as nn

from .captioner import Captioner
from .reward import RewardCriterion, get_ciderd_scorer, get_self_critical_reward
from .train import FusedClampAdam


def get_cls_reward(sample_captions, sample_masks, greedy_captions, greedy_masks, senti_labels, sent_senti_cls):
    """self_critical/utils.py:120-151 on device tensors: 1[argmax(classifier(sample)) == label] * the classifier's
    per-word weights, zero-padded to the caption length. Returns a float tensor [B, T].
    A classifier with ``forward_device`` (sent_senti_cls.SentenceSentimentClassifier) takes the lengths as a DEVICE tensor
    and returns weights already padded to T: no host round trip; any other module (the reference's class) gets the
    reference's python list of lengths."""
    training = sent_senti_cls.training
    T = sample_captions.shape[1]
    sent_senti_cls.eval()
    with torch.no_grad():
        if hasattr(sent_senti_cls, "forward_device"):
            preds, att_w = sent_senti_cls.forward_device(sample_captions, sample_masks.sum(dim=-1).to(torch.int32))
        else:
            lens = [int(x) for x in sample_masks.sum(dim=-1).to(torch.int).tolist()]
            preds, att_w = sent_senti_cls(sample_captions, lens)
        hit = (preds.softmax(dim=-1).argmax(dim=-1) == senti_labels).to(att_w.dtype).unsqueeze(1)
        scores = hit * att_w
    sent_senti_cls.train(training)
    return torch.nn.functional.pad(scores, (0, T - scores.shape[1])).float()


class Detector(nn.Module):
    def __init__(self, idx2word, max_seq_len, sentiment_categories, lrs, settings, senti_detector=None, sent_senti_cls=None,
                 precision="bf16x3"):
        super().__init__()
        self.idx2word = idx2word
        self.pad_id = idx2word.index("<PAD>")
        self.max_seq_len = max_seq_len
        self.captioner = Captioner(idx2word, sentiment_categories, settings, precision=precision)
        self.senti_detector = senti_detector
        self.sent_senti_cls = sent_senti_cls
        for mod in (self.senti_detector, self.sent_senti_cls):
            if mod is not None:
                mod.eval()
        self._lr = lrs["cap_lr"]
        self.cap_optim = None  # FusedClampAdam, created on first training use (the parameters must be on the GPU by then)
        _, self.cap_xe_crit, self.cap_da_crit = self.captioner.get_optim_criterion(self._lr)
        self.cap_rl_crit = RewardCriterion()
        self.cls_flag = 0.4
        self.seq_flag = 1.0
        self.senti_threshold = 0.7
        self.grad_clip = 0.1  # clip_gradient's default (models/decoder.py:14)
        self.sample_noise = None  # tests: Gumbel noise [T, B, V] for the sampled pass (forward_rl(noise=...))

    def set_ciderd_scorer(self, captions):
        self.ciderd_scorer = get_ciderd_scorer(captions, self.captioner.sos_id, self.captioner.eos_id,
                                               device=self.captioner._device())

    def set_sentiment_words(self, sentiment_words):
        self.sentiment_words = sentiment_words

    def set_lms(self, lms):
        self.lms = lms

    def _optim(self):
        if self.cap_optim is None:
            self.cap_optim = FusedClampAdam(self.captioner, lr=self._lr, grad_clip=self.grad_clip)
        return self.cap_optim

    def forward(self, data, data_type, training):
        """models/decoder.py:52-180. ``data`` = (caption batches, senti-corpus batches), any iterables with len()."""
        self.captioner.train(training)
        # every loss / reward stays a DEVICE scalar inside the loop and the sums come back in ONE stacked D2H copy after it
        # (the reference pulls seven floats per iteration, models/decoder.py:88-160)
        acc = defaultdict(lambda: 0.0)
        device = next(self.parameters()).device
        collect, self.captioner.collect_attention_weights = self.captioner.collect_attention_weights, False  # no consumer
        if training:
            seq2seq_data = iter(data[1])
        caption_data = iter(data[0])
        for _ in range(min(500, len(data[0]))):
            item = next(caption_data)
            if data_type == "fact":
                fns, fc_feats, att_feats, (caps_tensor, lengths), cpts_tensor, sentis_tensor, ground_truth = item
                caps_tensor = caps_tensor.to(device)
            elif data_type == "senti":
                fns, fc_feats, att_feats, cpts_tensor, sentis_tensor, senti_labels = item
                senti_labels = senti_labels.to(device)
            else:
                raise Exception("data_type(%s) is wrong!" % data_type)
            fc_feats, att_feats = fc_feats.to(device), att_feats.to(device)
            cpts_tensor, sentis_tensor = cpts_tensor.to(device), sentis_tensor.to(device)

            if data_type == "fact" or not training:
                if self.senti_detector is None:
                    raise RuntimeError("Detector: 'fact' batches (and evaluation) need the image sentiment detector: pass "
                                       "senti_detector= (e.g. the reference's SentimentDetector)")
                if hasattr(self.senti_detector, "sample_labels"):  # device labels only: no D2H for the sentiment names
                    senti_labels = self.senti_detector.sample_labels(att_feats, self.senti_threshold)
                else:
                    senti_labels, _, _, _ = self.senti_detector.sample(att_feats, self.senti_threshold)
                senti_labels = senti_labels.detach()

            sample_captions, sample_logprobs, seq_masks = self.captioner(
                fc_feats, att_feats, cpts_tensor, sentis_tensor, senti_labels, self.max_seq_len, sample_max=0, mode="rl",
                **({"noise": self.sample_noise} if self.sample_noise is not None else {}))
            da_loss = self.cap_da_crit(self.captioner.cpt_feats, self.captioner.fc_feats.detach())
            acc["da_loss"] = acc["da_loss"] + da_loss.detach()

            self.captioner.eval()
            with torch.no_grad():
                greedy_captions, _, greedy_masks = self.captioner(
                    fc_feats, att_feats, cpts_tensor, sentis_tensor, senti_labels, self.max_seq_len, sample_max=1, mode="rl")
            self.captioner.train(training)
            self.last_captions = (sample_captions, greedy_captions)  # device tensors of the last iteration (inspection)

            if data_type == "fact":
                fact_reward = get_self_critical_reward(sample_captions, greedy_captions, fns, ground_truth,
                                                       self.captioner.sos_id, self.captioner.eos_id, self.ciderd_scorer,
                                                       as_tensor=True).float()
                acc["fact_reward"] = acc["fact_reward"] + fact_reward[:, 0].mean()
            else:
                fact_reward = 0

            if self.sent_senti_cls is not None:
                cls_reward = get_cls_reward(sample_captions, seq_masks, greedy_captions, greedy_masks, senti_labels,
                                            self.sent_senti_cls)
                acc["cls_reward"] = acc["cls_reward"] + cls_reward.mean(-1).mean(-1)
                rewards = fact_reward + self.cls_flag * cls_reward
            else:
                rewards = fact_reward + torch.zeros_like(seq_masks)
            acc["all_rewards"] = acc["all_rewards"] + rewards.mean(-1).mean(-1)
            cap_loss = self.cap_rl_crit(sample_logprobs, seq_masks, rewards)
            acc["cap_loss"] = acc["cap_loss"] + cap_loss.detach()

            xe_loss = 0.0
            if data_type == "fact":
                if self.sent_senti_cls is None:
                    raise RuntimeError("Detector: 'fact' batches need the sentence sentiment classifier for the XE labels: "
                                       "pass sent_senti_cls= (e.g. the reference's SentenceSentimentClassifier)")
                with torch.no_grad():
                    xe_senti_labels, _ = self.sent_senti_cls(caps_tensor[:, 1:], lengths)
                    xe_senti_labels = xe_senti_labels.softmax(dim=-1).argmax(dim=-1).detach()
                pred = self.captioner(fc_feats, att_feats, cpts_tensor, caps_tensor, xe_senti_labels, ss_prob=0.5, mode="xe")
                xe_loss = self.cap_xe_crit(pred, caps_tensor[:, 1:], lengths)
                acc["xe_loss"] = acc["xe_loss"] + xe_loss.detach()

            seq2seq_loss = 0.0
            if training:
                try:
                    (s_caps, s_lengths), s_cpts, s_sentis, s_labels = next(seq2seq_data)
                except StopIteration:
                    seq2seq_data = iter(data[1])
                    (s_caps, s_lengths), s_cpts, s_sentis, s_labels = next(seq2seq_data)
                s_caps, s_cpts = s_caps.to(device), s_cpts.to(device)
                s_sentis, s_labels = s_sentis.to(device), s_labels.to(device)
                pred = self.captioner(s_caps, s_cpts, s_sentis, s_labels, ss_prob=0.25, mode="seq2seq")
                seq2seq_loss = self.seq_flag * self.cap_xe_crit(pred, s_caps[:, 1:], s_lengths)
                acc["seq2seq_loss"] = acc["seq2seq_loss"] + seq2seq_loss.detach()

            cap_loss = cap_loss + xe_loss + da_loss + seq2seq_loss
            if training:
                optim = self._optim()
                optim.zero_grad()
                optim.begin_iteration(2 + (data_type == "fact"))  # sampled re-score + seq2seq (+ xe) backward nodes
                cap_loss.backward()
                optim.step()  # clip_gradient (+-0.1) and Adam in one kernel

        self.captioner.collect_attention_weights = collect
        all_losses = defaultdict(float)
        if acc:
            keys = list(acc)
            vals = torch.stack([torch.as_tensor(acc[k], dtype=torch.float32, device=device).reshape(()) for k in keys]).tolist()
            for k, v in zip(keys, vals):
                all_losses[k] = v / len(data)  # the reference divides by len(data), a 2-tuple (decoder.py:178-179); kept
        return all_losses

    def sample(self, fc_feats, att_feats, sentis_tensor, beam_size=3, decoding_constraint=1):
        """models/decoder.py:182-192: detect the image's sentiment, then beam-search a caption conditioned on it."""
        self.eval()
        if self.senti_detector is None:
            raise RuntimeError("Detector.sample needs senti_detector=")
        att_feats = att_feats.unsqueeze(0)
        senti_label, _, det_img_sentis, _ = self.senti_detector.sample(att_feats, self.senti_threshold)
        captions, _ = self.captioner.sample(fc_feats, att_feats, sentis_tensor, senti_label, beam_size, decoding_constraint,
                                            self.max_seq_len)
        return captions, det_img_sentis
