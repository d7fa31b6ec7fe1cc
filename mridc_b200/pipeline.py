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

    One batch is always in flight on a private copy stream; the consumer's stream waits on the upload's event, and the
    device tensors are tied to the consumer's stream (``record_stream``) so the caching allocator cannot recycle them
    while kernels still read them.  Non-tensor values are passed through untouched.
    """

    def __init__(self, batches: Iterable[Dict[str, object]], device):
        self.batches = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPrefetcher: device must be a CUDA device (there is no CPU path)")

    def _upload(self, batch, stream):
        with torch.cuda.stream(stream):
            dev = {k: (v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
            ev = torch.cuda.Event()
            ev.record(stream)
        return dev, ev

    def __iter__(self) -> Iterator[Dict[str, object]]:
        stream = torch.cuda.Stream(self.device)
        it = iter(self.batches)
        try:
            pending = self._upload(next(it), stream)
        except StopIteration:
            return
        while pending is not None:
            cur, ev = pending
            consumer = torch.cuda.current_stream(self.device)
            consumer.wait_event(ev)
            for v in cur.values():
                if isinstance(v, torch.Tensor):
                    v.record_stream(consumer)
            try:
                pending = self._upload(next(it), stream)
            except StopIteration:
                pending = None
            yield cur
