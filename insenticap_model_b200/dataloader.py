"""Input pipeline of the path: the reference's datasets, collate functions and loader
factories (/root/reference/dataloader.py) over flat feature shards instead of per-item HDF5 opens.

* ``FeatureShard`` — one mmap'ed file of fixed-size records (fc_feats [D] + att_feats [L][D] per image, fp32 or bf16),
  read by the native ``isc_shard_*`` calls of libisc_b200.so; ``gather`` copies a batch's records into PINNED staging
  memory with a few host threads. ``FeatureShard.write`` converts in-memory arrays (e.g. read from the reference's
  ``*.h5`` files where h5py exists) once.
* ``create_collate_fn(name, ...)`` — the reference's seven collate functions (same names, same returned tuples, same
  ordering rules: captions sorted by length descending with Python's stable sort, ids padded / truncated to
  ``num_concepts`` / ``num_sentiments``, ``lengths - 1`` ...). Items may carry numpy feature arrays, as the reference's
  datasets return them, or lazy ``(shard, index)`` references that the collate function resolves with ONE batched gather.
* the seven ``Dataset`` classes and ``get_*_dataloader`` factories with the reference's signatures (feature arguments
  are shard paths or ``FeatureShard`` objects), and ``DevicePrefetcher`` which keeps ``depth`` batches in flight: the
  host gather of batch i+1 and its H2D copies (copy stream, pinned source) overlap the decode of batch i.
"""
from __future__ import annotations

import ctypes as C
import math
import queue
import random
import threading

import numpy as np
import torch
from torch.utils import data

from . import _lib

_DTYPES = {"fp32": 0, "bf16": 1, "fp16": 2}


class FeatureShard:
    """A read-only feature shard (include/isc.h: isc_shard_*). ``shard[fn]`` -> (fc, att) numpy arrays like the
    reference's ``f_fc[fn][:]`` / ``f_att[fn][:]``; ``gather(indices)`` is the batched form."""

    def __init__(self, path):
        self.path = str(path)
        lib = _lib.load()
        h = C.c_void_p()
        _lib.check(lib.isc_shard_open(self.path.encode(), C.byref(h)), "isc_shard_open")
        self._h = h
        n, d, l, dt = C.c_int64(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(lib.isc_shard_info(h, C.byref(n), C.byref(d), C.byref(l), C.byref(dt)), "isc_shard_info")
        self.n_images, self.feat_dim, self.n_regions = n.value, d.value, l.value
        self.dtype = {0: torch.float32, 1: torch.bfloat16, 2: torch.float16}[dt.value]
        self.direct_device = None  # set by pin(): batches are then DMA'd from the page cache, no staging copy
        g = math.isqrt(self.n_regions)
        self.att_shape = (g, g, self.feat_dim) if g * g == self.n_regions else (self.n_regions, self.feat_dim)

    @staticmethod
    def write(path, names, fc_feats, att_feats, dtype="fp32"):
        """names: list[str]; fc_feats [n, D], att_feats [n, ..., D] float32 arrays/tensors."""
        fc = np.ascontiguousarray(np.asarray(fc_feats, dtype=np.float32))
        att = np.ascontiguousarray(np.asarray(att_feats, dtype=np.float32))
        n, d = fc.shape
        if att.shape[0] != n or att.shape[-1] != d or len(names) != n or len(set(names)) != n:
            raise ValueError("write: names / fc_feats / att_feats disagree (or names repeat)")
        arr = (C.c_char_p * n)(*[str(s).encode() for s in names])
        _lib.check(_lib.load().isc_shard_write(str(path).encode(), _DTYPES[dtype], d, att[0].size // d, n, arr,
                                               fc.ctypes.data, att.ctypes.data), "isc_shard_write")
        return path

    def close(self):
        if getattr(self, "_h", None) and _lib is not None:  # (module globals are gone at interpreter shutdown)
            _lib.load().isc_shard_close(self._h)
            self._h = None

    __del__ = close

    def __len__(self):
        return self.n_images

    def index(self, fn):
        i = _lib.load().isc_shard_find(self._h, str(fn).encode())
        if i < 0:
            raise KeyError(fn)
        return i

    def __contains__(self, fn):
        return _lib.load().isc_shard_find(self._h, str(fn).encode()) >= 0

    def name(self, i):
        s = _lib.load().isc_shard_name(self._h, int(i))
        if s is None:
            raise IndexError(i)
        return s.decode()

    def gather(self, indices, want_fc=True, want_att=True, threads=16, pin=None):
        """-> (fc [n, D] | None, att [n, *att_shape] | None) torch tensors in the shard's dtype; pinned when a GPU is
        present (``pin`` overrides)."""
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int64))
        n = idx.shape[0]
        if pin is None:
            # a DataLoader worker is a forked child: CUDA (and so cudaHostAlloc) cannot be initialised there once the
            # parent has used it; workers gather into pageable memory and the loader's pin_memory thread / the consumer
            # page-locks it
            pin = torch.cuda.is_available() and torch.utils.data.get_worker_info() is None
        fc = torch.empty((n, self.feat_dim), dtype=self.dtype, pin_memory=pin) if want_fc else None
        att = torch.empty((n,) + self.att_shape, dtype=self.dtype, pin_memory=pin) if want_att else None
        if n == 0:
            return fc, att
        _lib.check(_lib.load().isc_shard_gather(self._h, idx.ctypes.data, n, fc.data_ptr() if want_fc else None,
                                                att.data_ptr() if want_att else None, int(threads)), "isc_shard_gather")
        return fc, att

    def pin(self, device="cuda:0"):
        """Page-lock the records (isc_shard_pin: the mapping in place, or a page-locked in-RAM copy of the file) so that
        batches go host -> HBM without the per-batch staging copy; lazy dataset items then collate straight into CUDA
        tensors on ``device`` (current stream). For shards that fit in host RAM. Returns False, leaving the staged path
        in place, if the memory cannot be page-locked."""
        dev = torch.device(device)
        with torch.cuda.device(dev):
            if _lib.load().isc_shard_pin(self._h) != 0:
                return False
        self.direct_device = dev
        return True

    def copy_to_device(self, indices, want_fc=True, want_att=True):
        """-> (fc, att) CUDA tensors on ``direct_device``, copies queued on its current stream (needs pin())."""
        dev = self.direct_device
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int64))
        n = idx.shape[0]
        fc = torch.empty((n, self.feat_dim), dtype=self.dtype, device=dev) if want_fc else None
        att = torch.empty((n,) + self.att_shape, dtype=self.dtype, device=dev) if want_att else None
        if n:
            with torch.cuda.device(dev):
                _lib.check(_lib.load().isc_shard_copy_to_device(self._h, idx.ctypes.data, n, fc.data_ptr() if want_fc else None,
                                                                att.data_ptr() if want_att else None, _lib.stream_ptr(dev)),
                           "isc_shard_copy_to_device")
        return fc, att

    def __getitem__(self, fn):
        fc, att = self.gather([self.index(fn)], threads=1, pin=False)
        if self.dtype != torch.float32:
            fc, att = fc.float(), att.float()
        return fc[0].numpy(), att[0].numpy()


def _shard(x):
    return x if isinstance(x, FeatureShard) or x is None else FeatureShard(x)


class _Lazy:
    """A feature still in its shard: (shard, record index, which = 'fc' | 'att'). Resolved per batch by the collate."""
    __slots__ = ("shard", "idx", "which")

    def __init__(self, shard, idx, which):
        self.shard, self.idx, self.which = shard, idx, which


class _Deferred:
    """Features of a page-locked shard whose copies have not been issued yet: (shard, record indices, 'fc' | 'att').
    ``DevicePrefetcher`` asks the collate for these (thread-local switch) so that the thread that runs the dataset / collate
    Python is not the one that sits inside the per-record cudaMemcpyAsync loop for the whole DMA time."""
    __slots__ = ("shard", "idx", "which")

    def __init__(self, shard, idx, which):
        self.shard, self.idx, self.which = shard, idx, which

    def resolve(self):
        fc, att = self.shard.copy_to_device(self.idx, want_fc=self.which == "fc", want_att=self.which == "att")
        return fc if self.which == "fc" else att


_tls = threading.local()


def _stack_features(feats):
    """The reference's ``torch.FloatTensor(np.array(feats))``; lazy references become one batched, threaded gather."""
    if feats and isinstance(feats[0], _Lazy):
        sh, which = feats[0].shard, feats[0].which
        # the direct path issues CUDA copies: only in the process that owns the CUDA context (not in a forked worker)
        if sh.direct_device is not None and torch.utils.data.get_worker_info() is None:
            d = _Deferred(sh, [f.idx for f in feats], which)
            return d if getattr(_tls, "defer_copies", False) else d.resolve()
        fc, att = sh.gather([f.idx for f in feats], want_fc=which == "fc", want_att=which == "att")
        return fc if which == "fc" else att
    return torch.from_numpy(np.array(feats, dtype=np.float32))


def _pad_ids(rows, width, pad_index):
    padded = []
    for r in rows:
        r = [int(x) for x in r[:width]]
        padded.append(r + [pad_index] * (width - len(r)))
    return torch.tensor(padded, dtype=torch.long).reshape(len(rows), width)


def _pad_captions(caps, max_seq_len, pad_index):
    """caps sorted by length descending -> (LongTensor [n, min(len(caps[0]), max_seq_len)], clipped lengths)."""
    lengths = [min(len(c), max_seq_len) for c in caps]
    return _pad_ids([c[:l] for c, l in zip(caps, lengths)], lengths[0], pad_index), lengths


def create_collate_fn(name, pad_index=0, max_seq_len=17, num_concepts=5, num_sentiments=10):
    """dataloader.py:9-151. Returns the collate function registered under ``name``."""

    def by_len(items, pos):  # stable, longest caption first
        return sorted(items, key=lambda it: len(it[pos]), reverse=True)

    def caption(batch):  # :11-34 — one row per (image, caption) pair
        rows = by_len([(fn, fc, att, cap, cpts) for fn, fc, att, caps, cpts in batch for cap in caps], 3)
        fns, fcs, atts, caps, cpts = zip(*rows)
        caps_tensor, lengths = _pad_captions(caps, max_seq_len, pad_index)
        return (fns, _stack_features(fcs), _stack_features(atts), (caps_tensor, [l - 1 for l in lengths]),
                _pad_ids(cpts, num_concepts, pad_index))

    def senti_corpus_with_sentis(batch):  # :36-58
        caps, cpts, sentis, senti_ids = zip(*by_len(batch, 0))
        caps_tensor, lengths = _pad_captions(caps, max_seq_len, pad_index)
        return ((caps_tensor, [l - 1 for l in lengths]), _pad_ids(cpts, num_concepts, pad_index),
                _pad_ids(sentis, num_sentiments, pad_index), torch.as_tensor(np.array(senti_ids), dtype=torch.long))

    def rl_fact(batch):  # :60-91 — one randomly drawn caption per image, all of them as ground truth
        ground_truth, rows = {}, []
        for fn, caps, fc, att, cpts, sentis in batch:
            ground_truth[fn] = [c[:max_seq_len] for c in caps]
            rows.append((fn, random.sample(caps, 1)[0], fc, att, cpts, sentis))
        fns, caps, fcs, atts, cpts, sentis = zip(*by_len(rows, 1))
        caps_tensor, lengths = _pad_captions(caps, max_seq_len, pad_index)
        return (fns, _stack_features(fcs), _stack_features(atts), (caps_tensor, [l - 1 for l in lengths]),
                _pad_ids(cpts, num_concepts, pad_index), _pad_ids(sentis, num_sentiments, pad_index), ground_truth)

    def rl_senti(batch):  # :93-109
        fns, fcs, atts, cpts, sentis, labels = zip(*batch)
        return (fns, _stack_features(fcs), _stack_features(atts), _pad_ids(cpts, num_concepts, pad_index),
                _pad_ids(sentis, num_sentiments, pad_index), torch.as_tensor(np.array(labels), dtype=torch.long))

    def concept(batch):  # :111-115
        fns, fcs, cpts = zip(*batch)
        return fns, _stack_features(fcs), torch.as_tensor(np.array(cpts), dtype=torch.long)

    def senti_image(batch):  # :117-121
        fns, atts, labels = zip(*batch)
        return fns, _stack_features(atts), torch.as_tensor(np.array(labels), dtype=torch.long)

    def senti_sents(batch):  # :123-135 — lengths are NOT reduced by one here
        sentis, caps = zip(*by_len(batch, 1))
        caps_tensor, lengths = _pad_captions(caps, max_seq_len, pad_index)
        return torch.as_tensor(np.array(sentis), dtype=torch.long), (caps_tensor, lengths)

    fns = {"caption": caption, "senti_sents": senti_sents, "concept": concept, "senti_image": senti_image,
           "rl_fact": rl_fact, "rl_senti": rl_senti, "senti_corpus_with_sentis": senti_corpus_with_sentis}
    return fns.get(name)  # the reference returns None for an unknown name too


class _ShardDataset(data.Dataset):
    """Feature access shared by the image datasets: ``lazy`` (default) hands the collate function a reference into the
    shard, so a batch costs one threaded gather; ``lazy=False`` returns numpy arrays item by item, like the reference."""

    def _open(self, fc_feats, att_feats, lazy):
        self.fc_shard, self.att_shard, self.lazy = _shard(fc_feats), _shard(att_feats), lazy

    def _fc(self, fn):
        sh = self.fc_shard
        return _Lazy(sh, sh.index(fn), "fc") if self.lazy else sh[fn][0]

    def _att(self, fn):
        sh = self.att_shard
        return _Lazy(sh, sh.index(fn), "att") if self.lazy else sh[fn][1]


class SCSDataset(data.Dataset):  # dataloader.py:154-163
    def __init__(self, senti_corpus_with_sentis):
        self.senti_corpus_with_sentis = senti_corpus_with_sentis

    def __getitem__(self, index):
        cap, cpts, sentis, senti_id = self.senti_corpus_with_sentis[index]
        return cap, cpts, sentis, senti_id

    def __len__(self):
        return len(self.senti_corpus_with_sentis)


class CaptionDataset(_ShardDataset):  # :166-183
    def __init__(self, fc_feats, att_feats, img_captions, img_det_concepts, lazy=True):
        self._open(fc_feats, att_feats, lazy)
        self.captions = list(img_captions.items())
        self.det_concepts = img_det_concepts

    def __getitem__(self, index):
        fn, caps = self.captions[index]
        return fn, self._fc(fn), self._att(fn), caps, self.det_concepts[fn]

    def __len__(self):
        return len(self.captions)


class RLFactDataset(_ShardDataset):  # :186-206
    def __init__(self, fc_feats, att_feats, img_captions, img_det_concepts, img_det_sentiments, lazy=True):
        self._open(fc_feats, att_feats, lazy)
        self.captions = list(img_captions.items())
        self.det_concepts = img_det_concepts
        self.det_sentiments = img_det_sentiments

    def __getitem__(self, index):
        fn, caps = self.captions[index]
        return fn, caps, self._fc(fn), self._att(fn), self.det_concepts[fn], self.det_sentiments[fn]

    def __len__(self):
        return len(self.captions)


class RLSentiDataset(_ShardDataset):  # :209-229
    def __init__(self, fc_feats, att_feats, img_det_concepts, img_det_sentiments, img_senti_labels, lazy=True):
        self._open(fc_feats, att_feats, lazy)
        self.det_concepts = img_det_concepts
        self.det_sentiments = img_det_sentiments
        self.img_senti_labels = img_senti_labels

    def __getitem__(self, index):
        fn, senti_label = self.img_senti_labels[index]
        return fn, self._fc(fn), self._att(fn), self.det_concepts[fn], self.det_sentiments[fn], senti_label

    def __len__(self):
        return len(self.img_senti_labels)


class ConceptDataset(_ShardDataset):  # :232-247
    def __init__(self, fc_feats, img_concepts, num_cpts, lazy=True):
        self._open(fc_feats, None, lazy)
        self.concepts = list(img_concepts.items())
        self.num_cpts = num_cpts

    def __getitem__(self, index):
        fn, cpts_idx = self.concepts[index]
        cpts = np.zeros(self.num_cpts, dtype=np.int16)
        cpts[cpts_idx] = 1
        return fn, self._fc(fn), cpts

    def __len__(self):
        return len(self.concepts)


class SentiImageDataset(_ShardDataset):  # :250-262
    def __init__(self, senti_att_feats, img_senti_labels, lazy=True):
        self._open(None, senti_att_feats, lazy)
        self.img_senti_labels = img_senti_labels

    def __getitem__(self, index):
        fn, senti_label = self.img_senti_labels[index]
        return fn, self._att(fn), senti_label

    def __len__(self):
        return len(self.img_senti_labels)


class SentiSentDataset(data.Dataset):  # :265-274
    def __init__(self, senti_sentences):
        self.senti_sentences = senti_sentences

    def __getitem__(self, index):
        senti, sent = self.senti_sentences[index]
        return senti, np.array(sent)

    def __len__(self):
        return len(self.senti_sentences)


def _loader(dataset, batch_size, num_workers, shuffle, collate):
    return data.DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, collate_fn=collate)


def get_caption_dataloader(fc_feats, att_feats, img_captions, img_det_concepts, pad_index, max_seq_len, num_concepts,
                           batch_size, num_workers=0, shuffle=True):  # :277-288
    return _loader(CaptionDataset(fc_feats, att_feats, img_captions, img_det_concepts), batch_size, num_workers, shuffle,
                   create_collate_fn("caption", pad_index, max_seq_len + 1, num_concepts))


def get_senti_corpus_with_sentis_dataloader(senti_corpus_with_sentis, pad_index, max_seq_len, num_concepts, num_sentiments,
                                            batch_size, num_workers=0, shuffle=True):  # :291-303
    return _loader(SCSDataset(senti_corpus_with_sentis), batch_size, num_workers, shuffle,
                   create_collate_fn("senti_corpus_with_sentis", pad_index, max_seq_len + 1, num_concepts=num_concepts,
                                     num_sentiments=num_sentiments))


def get_rl_fact_dataloader(fc_feats, att_feats, img_captions, img_det_concepts, img_det_sentiments, pad_index, max_seq_len,
                           num_concepts, num_sentiments, batch_size, num_workers=0, shuffle=True):  # :306-320
    return _loader(RLFactDataset(fc_feats, att_feats, img_captions, img_det_concepts, img_det_sentiments), batch_size,
                   num_workers, shuffle, create_collate_fn("rl_fact", pad_index=pad_index, max_seq_len=max_seq_len + 1,
                                                           num_concepts=num_concepts, num_sentiments=num_sentiments))


def get_rl_senti_dataloader(fc_feats, att_feats, img_det_concepts, img_det_sentiments, img_senti_labels, pad_index,
                            num_concepts, num_sentiments, batch_size, num_workers=0, shuffle=True):  # :323-336
    return _loader(RLSentiDataset(fc_feats, att_feats, img_det_concepts, img_det_sentiments, img_senti_labels), batch_size,
                   num_workers, shuffle, create_collate_fn("rl_senti", pad_index=pad_index, num_concepts=num_concepts,
                                                           num_sentiments=num_sentiments))


def get_concept_dataloader(fc_feats, img_concepts, num_cpts, batch_size, num_workers=0, shuffle=True):  # :339-347
    return _loader(ConceptDataset(fc_feats, img_concepts, num_cpts), batch_size, num_workers, shuffle,
                   create_collate_fn("concept"))


def get_senti_image_dataloader(senti_att_feats, img_senti_labels, batch_size, num_workers=0, shuffle=True):  # :350-357
    return _loader(SentiImageDataset(senti_att_feats, img_senti_labels), batch_size, num_workers, shuffle,
                   create_collate_fn("senti_image"))


def get_senti_sents_dataloader(senti_sentences, pad_index, max_seq_len, batch_size=80, num_workers=2, shuffle=True):  # :360-370
    return _loader(SentiSentDataset(senti_sentences), batch_size, num_workers, shuffle,
                   create_collate_fn("senti_sents", pad_index=pad_index, max_seq_len=max_seq_len))


def _to_device(x, dev):
    if isinstance(x, _Deferred):
        return x.resolve()  # copies queued on the current stream of the shard's device
    if torch.is_tensor(x):
        return x.to(dev, non_blocking=True)
    if isinstance(x, tuple) and not (x and isinstance(x[0], str)):
        return tuple(_to_device(y, dev) for y in x)
    return x  # names, lengths lists, ground-truth dicts stay on the host


class DevicePrefetcher:
    """Iterate ``loader`` with ``depth`` batches in flight. Two worker threads: one runs the loader (dataset reads + collate,
    i.e. the shard gather into pinned memory — or, for a page-locked shard, just the record indices), the other issues the
    H2D copies on a private copy stream (a page-locked shard's per-record copies keep their issuing thread busy for the
    whole DMA time: with one thread for both, the Python of batch i+1 waited for the copies of batch i). The consumer's
    stream waits on each batch's event, so the copies of batch i+1 overlap whatever the consumer does with batch i."""

    def __init__(self, loader, device, depth=2):
        self.loader, self.device, self.depth = loader, torch.device(device), int(depth)

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        # collated host batches (features: pinned tensors or _Deferred). One slot: the collate thread only has to stay one
        # batch ahead of the copy thread, and every batch in flight holds a pinned staging block whose first allocation
        # (cudaHostAlloc of 0.4-0.8 GB) stalls the pipeline for 0.1-0.3 s
        host_q = queue.Queue(maxsize=1)
        q = queue.Queue(maxsize=self.depth)       # (device batch, host batch, copy event)
        stream = torch.cuda.Stream(self.device)
        stop = threading.Event()

        def put(qq, item):
            while not stop.is_set():
                try:
                    qq.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def collate_work():
            try:
                _tls.defer_copies = True  # page-locked shards: leave the copies to copy_work
                for batch in self.loader:
                    if not put(host_q, batch):
                        return
                put(host_q, None)
            except BaseException as e:  # surface loader errors in the consumer
                put(host_q, e)

        def copy_work():
            try:
                while not stop.is_set():
                    try:
                        batch = host_q.get(timeout=0.1)
                    except queue.Empty:
                        continue
                    if batch is None or isinstance(batch, BaseException):
                        put(q, batch)
                        return
                    with torch.cuda.stream(stream):
                        dev_batch = _to_device(batch, self.device)
                        ev = torch.cuda.Event()
                        ev.record(stream)
                    if not put(q, (dev_batch, batch, ev)):  # the host batch stays alive until its copy is done
                        return
            except BaseException as e:
                put(q, e)

        threads = [threading.Thread(target=collate_work, daemon=True), threading.Thread(target=copy_work, daemon=True)]
        for th in threads:
            th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                dev_batch, host_batch, ev = item
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                for t in _tensors(dev_batch):
                    t.record_stream(cur)
                ev.synchronize()  # the pinned source may be reused once the copy has landed
                del host_batch
                yield dev_batch
        finally:
            stop.set()
            for th in threads:
                th.join(timeout=5)


def _tensors(x):
    if torch.is_tensor(x):
        yield x
    elif isinstance(x, tuple):
        for y in x:
            yield from _tensors(y)
