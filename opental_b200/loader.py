"""Background batch assembly + host->device prefetch for the uint8 window datasets  (SURVEY §7 hard part 7: feeding 8 GPUs).

The reference feeds the GPU through `DataLoader(num_workers=4, pin_memory=True)` of fp32 clips (train.py:349-353): worker
PROCESSES crop / flip / normalise / augment on the CPU and ship 28 MB per clip (plus the same again for the cut-paste clip).
Here a sample is the uint8 window as stored (9.6 MB) plus a few integers (opental_b200/dataset.py), so the host side only has
to copy windows out of the memory-mapped `.npy` files; at 340 clips/s per GPU that is still 3.3 GB/s per rank, which one Python
thread cannot assemble and a pageable `tensor.to(device)` cannot ship without stalling the stream.  `Prefetcher`:

  * `workers` THREADS (numpy's copies release the GIL) build whole batches, step s by one thread, straight into a ring of
    pre-allocated PINNED buffers — no `np.stack` temporary, no per-batch `cudaHostAlloc`;
  * randomness is per step — `random.Random` seeded by (seed, epoch, rank, step) — so batches do not depend on thread timing and
    equal `dataset.epoch_batches` of the same arguments;
  * the consumer enqueues step s's copies on a COPY STREAM when it is asked for step s, i.e. right after step s-1's kernels were
    enqueued, so the transfer runs under step s-1's compute; two device slots alternate, guarded by events exactly as in
    `bench.py`'s end-to-end loop (the pattern measured there: 339 vs 343 clips/s with resident inputs);
  * targets travel padded to a fixed number of slots ((tensor [B,G,3], mask [B,G]) — what a captured step graph wants).

Without a device (CPU tests) the same threads / ring / ordering run and host batches are yielded."""
from __future__ import annotations

import random as _random
import threading

import numpy as np
import torch

from .multisegment_loss import pad_targets


def step_indices(n_items: int, batch_size: int, epoch: int, *, rank: int = 0, world: int = 1, seed: int = 0, shuffle: bool = True):
    """Per step, the window indices of this rank: one shuffle shared by all ranks (seeded by seed and epoch), global batches of
    batch_size * world windows, drop_last (train.py:349-353), rank r takes the r-th slice."""
    order = list(range(n_items))
    if shuffle:
        _random.Random(1_000_003 * seed + epoch).shuffle(order)
    per_step = batch_size * world
    return [order[s * per_step + rank * batch_size: s * per_step + (rank + 1) * batch_size] for s in range(n_items // per_step)]


def step_rng(seed: int, epoch: int, rank: int, step: int) -> _random.Random:
    """The generator that draws crop / mirror / cut-paste decisions of one step's samples, in sample order."""
    return _random.Random(((1_000_003 * seed + epoch) * 4099 + rank + 1) * 1_000_033 + step)


class Prefetcher:
    def __init__(self, ds, batch_size: int, epoch: int, *, rank: int = 0, world: int = 1, seed: int = 0, device=None, workers: int = 4,
                 depth: int = 3, target_slots: int | None = None, crop_offsets: torch.Tensor | None = None, shuffle: bool = True,
                 ssl: bool = True):
        """target_slots: ground-truth slots of the padded targets; default = the largest number of segments of any window of the
        dataset (at least 8), so that one captured step graph serves the whole epoch."""
        self.ds, self.B, self.epoch, self.rank, self.seed = ds, batch_size, epoch, rank, seed
        self.steps = step_indices(len(ds), batch_size, epoch, rank=rank, world=world, seed=seed, shuffle=shuffle)
        self.device = torch.device(device) if device is not None else None
        self.cuda = self.device is not None and self.device.type == "cuda"
        self.workers, self.ring = max(1, workers), max(2, depth + 1)
        if target_slots is None:
            target_slots = max([8] + [len(w["annos"]) for w in getattr(ds, "training_list", [])])
        self.target_slots, self.crop_offsets, self.ssl = int(target_slots), crop_offsets, ssl
        self._host = [None] * self.ring          # ring of host slots (allocated lazily: the first sample tells the geometry)
        self._lock = threading.Condition()
        self._built: dict[int, dict] = {}        # step -> host batch (references ring slot step % ring)
        self._next = 0                           # next step a worker may take
        self._released = -1                      # highest step whose ring slot the consumer has handed back
        self._error: BaseException | None = None
        self._stop = False

    def __len__(self) -> int:
        return len(self.steps)

    # ------------------------------------------------------------------------------------------------ host side (worker threads)
    def _alloc_slot(self, frames_shape, T: int, score_rows: int) -> dict:
        pin = self.cuda
        mk = lambda *shape, dtype: torch.empty(*shape, dtype=dtype, pin_memory=pin)      # noqa: E731
        B, G = self.B, self.target_slots
        return dict(clips=mk(B, *frames_shape, dtype=torch.uint8), tgt=mk(B, G, 3, dtype=torch.float32), valid=mk(B, G, dtype=torch.bool),
                    scores=mk(B, score_rows, T, dtype=torch.float32), fmap=mk(B, T, dtype=torch.int32), ssl_t=mk(B, 3, 2, dtype=torch.float32),
                    crop=mk(B, 3, dtype=torch.int32))

    def _build(self, step: int) -> dict:
        rng = step_rng(self.seed, self.epoch, self.rank, step)
        samples = [self.ds.sample(i, rng) for i in self.steps[step]]
        slot_id = step % self.ring
        if self._host[slot_id] is None:
            s0 = samples[0]
            self._host[slot_id] = self._alloc_slot(s0["frames"].shape, s0["frames"].shape[0], s0["scores"].shape[0])
        h = self._host[slot_id]
        clips = h["clips"].numpy()
        for j, s in enumerate(samples):
            np.copyto(clips[j], s["frames"])                                  # the one big copy, GIL released
        tgt, valid = pad_targets([torch.from_numpy(s["target"]) for s in samples], None, slots=self.target_slots)
        h["tgt"].copy_(tgt), h["valid"].copy_(valid)
        h["scores"].copy_(torch.from_numpy(np.stack([s["scores"] for s in samples])))
        h["fmap"].copy_(torch.from_numpy(np.stack([s["frame_map"] for s in samples]).astype(np.int32)))
        placeholder = np.asarray([[0, 1], [1, 2], [2, 3]], dtype=np.float32)       # unused unless flags[0] (train.py:237)
        h["ssl_t"].copy_(torch.from_numpy(np.stack([s["ssl_target"][:3] if s["flag"] and len(s["ssl_target"]) >= 3 else placeholder
                                                    for s in samples])))
        h["crop"].copy_(torch.tensor([s["crop"] for s in samples], dtype=torch.int32))
        return dict(slot=h, flags=[s["flag"] for s in samples], step=step)

    def _worker(self) -> None:
        try:
            while True:
                with self._lock:
                    # a step may be built once the ring slot it writes is free: step - ring has been handed back
                    while not self._stop and (self._next >= len(self.steps) or self._next - self.ring > self._released):
                        if self._next >= len(self.steps):
                            return
                        self._lock.wait(0.05)
                    if self._stop:
                        return
                    step = self._next
                    self._next += 1
                batch = self._build(step)
                with self._lock:
                    self._built[step] = batch
                    self._lock.notify_all()
        except BaseException as e:  # noqa: BLE001 - handed to the consumer
            with self._lock:
                self._error = e
                self._lock.notify_all()

    # ------------------------------------------------------------------------------------------------ consumer
    def _take(self, step: int) -> dict:
        with self._lock:
            while step not in self._built and self._error is None:
                self._lock.wait(0.05)
            if self._error is not None:
                raise RuntimeError("batch assembly failed in a loader thread") from self._error
            return self._built.pop(step)

    def _release(self, step: int) -> None:
        with self._lock:
            self._released = max(self._released, step)
            self._lock.notify_all()

    def _as_batch(self, t: dict, flags) -> dict:
        b = dict(clips=t["clips"], targets=(t["tgt"], t["valid"]), scores=t["scores"], flags=flags, crop_offsets=t["crop"])
        if self.ssl:
            b.update(ssl_frame_map=t["fmap"], ssl_targets=list(t["ssl_t"].unbind(0)))
        return b

    def __iter__(self):
        threads = [threading.Thread(target=self._worker, daemon=True, name=f"otal-loader-{k}") for k in range(self.workers)]
        for th in threads:
            th.start()
        try:
            if not self.cuda:
                for s in range(len(self.steps)):
                    hb = self._take(s)
                    out = {k: v.clone() for k, v in hb["slot"].items()}          # the ring slot goes back to the workers
                    self._release(s)
                    if self.crop_offsets is not None:
                        self.crop_offsets.copy_(out["crop"])
                    yield self._as_batch(out, hb["flags"])
                return
            copy_stream = torch.cuda.Stream(self.device)
            dev = [None, None]                                                   # two device slots
            ready = [torch.cuda.Event(), torch.cuda.Event()]                     # copy of slot k finished
            consumed = [torch.cuda.Event(), torch.cuda.Event()]                  # kernels that read slot k were enqueued before this
            main = torch.cuda.current_stream(self.device)
            for e in consumed:
                e.record(main)
            for s in range(len(self.steps)):
                k = s % 2
                if s > 0:
                    consumed[1 - k].record(main)                                 # step s-1's kernels are in the stream by now
                hb = self._take(s)
                if dev[k] is None:
                    dev[k] = {name: torch.empty_like(v, device=self.device) for name, v in hb["slot"].items()}
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[k])                          # slot k was last read by step s-2
                    for name, v in hb["slot"].items():
                        dev[k][name].copy_(v, non_blocking=True)
                    ready[k].record(copy_stream)
                if s > 0:
                    ready[1 - k].synchronize()                                   # step s-1's pinned slot has been read: hand it back
                    self._release(s - 1)
                main.wait_event(ready[k])
                if self.crop_offsets is not None:
                    self.crop_offsets.copy_(dev[k]["crop"], non_blocking=True)   # stream-ordered after step s-1's ingest kernel
                yield self._as_batch(dev[k], hb["flags"])
        finally:
            with self._lock:
                self._stop = True
                self._lock.notify_all()
            for th in threads:
                th.join(timeout=5.0)
