"""The recognition step's batching contract on B200: the stand-in for the reference's in-process server.

  Bucket              <- src/processing/server.py:17-57   (32-px width buckets, right zero-padding, fixed batch)
  LocalServer         <- src/processing/server.py:60-145  (31 buckets (32k, 32k+32], filler crops, graph call,
                                                           strip -1, map through the charset)
  BatchLinePredictor  <- src/processing/linepredictor.py:11-36  (predict_batch(batch_name, img_list, logger) -> {i: text})

What is kept: the bucket boundaries, the padding value (uint8 0, i.e. -0.5 after preprocessing, not the
training pipeline's 0.0), the fixed batch size with zero-crop fillers that reuse widths[0], the strict
"more than batchsize" / "older than maxtime" release rule, and the output post-processing.
What is not: the multiprocessing.Manager queues and the forever loop with 0.1 s sleeps (product plumbing,
SURVEY.md section 2 rows 10-11 "boundary only").  LocalServer here is synchronous: submit() crops, then flush()
or poll(); `predict_batch` gives the BatchLinePredictor call directly.
Known reference quirk kept visible, not copied: a crop exactly 32 px wide matches no bucket and trips the
reference's assert (server.py:29,114); here it raises ValueError.
"""
import time

import numpy as np
import torch

from . import model as _model


class Bucket(object):
    """server.py:17-57."""

    def __init__(self, maxtime, batchsize, widthrange):
        self.maxtime = maxtime
        self.batchsize = batchsize
        self.widthrange = widthrange
        self.imgs = []
        self.widths = []
        self.infos = []
        self.oldesttime = None

    def addImgToBucket(self, clientid, imgid, imgtime, img):
        w = img.shape[1]
        if w > self.widthrange[0] and w <= self.widthrange[1]:
            if img.ndim == 3:
                img = img[:, :, 1]           # the reference keeps channel 1 of a colour crop (server.py:33-34)
            newimg = np.zeros((img.shape[0], self.widthrange[1]), np.uint8)   # cv2.copyMakeBorder(..., value=0)
            newimg[:, :w] = img
            self.imgs.append(newimg[:, :, np.newaxis])
            self.widths.append(w)
            self.infos.append((clientid, imgid))
            if self.oldesttime is None or imgtime < self.oldesttime:
                self.oldesttime = imgtime
            return True
        return False

    def getBatch(self, force=False):
        if len(self.imgs) == 0:
            return None
        if force or len(self.imgs) > self.batchsize or (time.time() - self.oldesttime) > self.maxtime:
            batch = np.array(self.imgs[:self.batchsize])
            widths = np.array(self.widths[:self.batchsize], np.int32)
            infos = self.infos[:self.batchsize]
            self.imgs = self.imgs[self.batchsize:]
            self.widths = self.widths[self.batchsize:]
            self.infos = self.infos[self.batchsize:]
            self.oldesttime = time.time()
            return infos, batch, widths
        return None


class LocalServer(object):
    """Synchronous version of server.py:60-145 around a `model.Model`."""

    def __init__(self, recognizer, bucket_size=32, bucket_max_time=0.5, device="cuda", shard=None):
        """shard = (rank, world): multi-GPU dispatch with one server process per GPU and no collective.  Every process is
        handed the same stream of crops; a crop's batch is known when it arrives (bucket k, the n-th crop of that bucket ->
        batch n // bucket_size), and batch (k, j) belongs to process (k + j) % world.  Whole BATCHES are dealt, so the
        fillers are the same as on one GPU (only a bucket's last batch is ever short) and neighbouring widths -- similar
        cost -- land on different GPUs.  A process keeps only its own crops; `take` returns only its own strings."""
        self.model = recognizer
        self.bucket_size = bucket_size
        self.device = torch.device(device)
        self.buckets = [Bucket(bucket_max_time, bucket_size, (w, w + 32)) for w in range(32, 1000, 32)]  # server.py:64-65
        self.rank, self.world = shard if shard is not None else (0, 1)
        self._seen = [0] * len(self.buckets)   # crops offered to each bucket so far, owned or not
        self.results = {}
        self._pending = []
        self.padded_pixels = 0
        self.real_pixels = 0

    def submit(self, clientid, imgid, img, imgtime=None):
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.shape[0] != 32:
            raise ValueError("crops must be uint8 with height 32 (pagepredictor2.py:102-105)")
        if self.world > 1:
            w = img.shape[1]
            k = (w - 1) // 32 - 1    # bucket k holds widths in (32k+32, 32k+64]
            if w <= 32 or k >= len(self.buckets):
                raise ValueError("crop width %d falls in no bucket: widths must be in (32, 1024]" % w)
            j = self._seen[k] // self.bucket_size
            self._seen[k] += 1
            if (k + j) % self.world != self.rank:
                return False
            self.buckets[k].addImgToBucket(clientid, imgid, time.time() if imgtime is None else imgtime, img)
            return True
        ok = sum(b.addImgToBucket(clientid, imgid, time.time() if imgtime is None else imgtime, img) for b in self.buckets)
        if ok != 1:
            raise ValueError("crop width %d falls in no bucket: widths must be in (32, 1024]" % img.shape[1])
        return True

    def _run_batch(self, infos, batch, widths):
        n = batch.shape[0]
        if n < self.bucket_size:   # server.py:125-130: zero crops, widths[0]
            batch = np.concatenate((batch, np.zeros((self.bucket_size - n,) + batch.shape[1:], np.uint8)))
            widths = np.concatenate((widths, np.full(self.bucket_size - n, widths[0], np.int32)))
            infos = infos + [("-1", "0")] * (self.bucket_size - n)
        self.padded_pixels += int(batch.shape[0] * batch.shape[2])
        self.real_pixels += int(widths[:n].sum())
        if hasattr(self.model, "recognize_async"):
            # host crops are copied into the recorded graph's input; the strings are collected one batch later, so the
            # host assembles the next batch while this one runs on the GPU
            self._pending.append((infos, self.model.recognize_async(torch.from_numpy(batch), torch.from_numpy(widths))))
            while len(self._pending) > 1:
                self._collect()
        else:
            self._store(infos, self.model.recognize(torch.from_numpy(batch), torch.from_numpy(widths)))

    def _collect(self):
        infos, handle = self._pending.pop(0)
        self._store(infos, handle.result())

    def _store(self, infos, texts):
        for (clientid, imgid), txt in zip(infos, texts):
            if clientid == "-1":
                continue
            self.results.setdefault(clientid, {})[imgid] = txt

    def _drain(self):
        while self._pending:
            self._collect()

    def poll(self):
        """One pass of the reference's loop body over the buckets (server.py:120-140)."""
        for bucket in self.buckets:
            b = bucket.getBatch()
            if b is not None:
                self._run_batch(*b)
        self._drain()

    def flush(self):
        """Release every bucket regardless of fill level or age."""
        for bucket in self.buckets:
            while True:
                b = bucket.getBatch(force=True)
                if b is None:
                    break
                self._run_batch(*b)
        self._drain()
        self._seen = [0] * len(self.buckets)   # every bucket is empty again: batch numbering restarts (the same on every process)

    def take(self, clientid):
        return self.results.pop(clientid, {})


class BatchLinePredictor(object):
    """linepredictor.py:11-36 against a LocalServer in the same process."""

    def __init__(self, server, logger=None):
        self.server = server
        self.clientid = str(id(self))

    def predict_batch(self, batch_name, img_list, logger=None):
        for i, img in enumerate(img_list):
            self.server.submit(self.clientid, batch_name + "_" + str(i), img)
        self.server.flush()
        got = self.server.take(self.clientid)
        if self.server.world > 1:   # this process's share; the caller gathers the dictionaries of all processes on the host
            return {i: got[batch_name + "_" + str(i)] for i in range(len(img_list)) if batch_name + "_" + str(i) in got}
        return {i: got[batch_name + "_" + str(i)] for i in range(len(img_list))}
