"""Host -> device staging for slice streams.

The reference feeds its models from a ``DataLoader`` with ``pin_memory`` and uploads each batch on the compute stream
(``reconstruction/models/base.py:700-712`` drives ``test_step`` batch by batch).  Here the upload of batch k+1 runs on a
copy stream while batch k is being reconstructed, so the PCIe/C2C transfer disappears behind the unrolled network.
"""
from typing import Dict, Iterable, Iterator

import torch

__all__ = ["HostPrefetcher"]


class HostPrefetcher:
    """Iterate over dictionaries of (pinned) host tensors, yielding the same dictionaries on ``device``.

    One batch is always in flight on a private copy stream; the consumer's stream waits on the upload's event.  The
    device tensors are two persistent buffer sets that alternate: a yielded batch is valid ONLY UNTIL THE NEXT ADVANCE of the
    iterator.  Requesting batch i + 1 enqueues the upload of batch i + 2 into batch i's buffers, ordered after everything
    the consumer's stream had been given before that advance -- work enqueued on batch i afterwards would race with the
    overwrite, so clone what must live longer.  Non-tensor values are passed through untouched.
    """

    def __init__(self, batches: Iterable[Dict[str, object]], device):
        self.batches = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPrefetcher: device must be a CUDA device (there is no CPU path)")

    def _upload(self, batch, stream, slot):
        bufs = self._bufs[slot]
        with torch.cuda.stream(stream):
            if self._done[slot] is not None:
                stream.wait_event(self._done[slot])  # the consumer has finished with the batch that lived in this slot
            dev = {}
            for k, v in batch.items():
                if isinstance(v, torch.Tensor):
                    dst = bufs.get(k)
                    if dst is None or dst.shape != v.shape or dst.dtype != v.dtype:
                        dst = bufs[k] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    dst.copy_(v, non_blocking=True)
                    dev[k] = dst
                else:
                    dev[k] = v
            ev = torch.cuda.Event()
            ev.record(stream)
        return dev, ev

    def __iter__(self) -> Iterator[Dict[str, object]]:
        # two persistent sets of device buffers (no allocator traffic in steady state): batch i lives in slot i % 2 and
        # is overwritten by batch i + 2 only after the consumer's work on batch i has been enqueued and has finished
        stream = torch.cuda.Stream(self.device)
        self._bufs = [{}, {}]
        self._done = [None, None]
        it = iter(self.batches)
        try:
            pending = self._upload(next(it), stream, 0)
        except StopIteration:
            return
        i = 0
        while pending is not None:
            cur, ev = pending
            consumer = torch.cuda.current_stream(self.device)
            consumer.wait_event(ev)
            try:
                pending = self._upload(next(it), stream, (i + 1) % 2)
            except StopIteration:
                pending = None
            yield cur
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            self._done[i % 2] = done
            i += 1
