"""Serving-side runtime for the S2M2 forward: one captured CUDA graph per (batch, H, W) plus a double-buffered host pipeline.

``S2M2.forward`` enqueues ~900 kernel launches; replaying them from a CUDA graph removes the host launch cost, and a second
stream moves the NEXT batch's images host→device and the PREVIOUS batch's maps device→host while the current batch computes.
Pairs are independent, so this is the whole data-parallel runtime: one ``PipelinedRunner`` per GPU / process.

    runner = PipelinedRunner(model, batch=8, height=1024, width=1216)
    for disp, occ, conf in runner.stream(batches):      # batches: iterable of (left, right) pinned host tensors (B,3,H,W)
        ...                                             # pinned host tensors, valid until two batches later

The reference's wrapper (``run_stereo_matching``, model_utils.py:50-95) does the same steps serially: ``.to(device)``, forward,
synchronise, read back.
"""
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

Maps = Tuple[torch.Tensor, torch.Tensor, torch.Tensor]


class PipelinedRunner:
    def __init__(self, model, batch: int, height: int, width: int, device: Optional[torch.device] = None,
                 input_dtype: torch.dtype = torch.float32, use_graph: bool = True):
        if input_dtype not in (torch.float32, torch.uint8):
            raise ValueError("images are float32 (what image_pad produces) or uint8")
        self.model = model
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (batch, 3, height, width)
        dev = self.device
        # static device buffers the graph reads / writes
        self.left = torch.zeros(self.shape, dtype=input_dtype, device=dev)
        self.right = torch.zeros(self.shape, dtype=input_dtype, device=dev)
        self.graph = None
        if hasattr(model, "auto_graph"):
            model.auto_graph = False      # this runner owns the captured graph (the module's own same-shape replay would hold a second pool)
        with torch.inference_mode():
            for _ in range(2):                        # packs the weights, builds tables, sets kernel attributes
                out = model(self.left, self.right)
            torch.cuda.synchronize(dev)
            if use_graph:
                del out
                torch.cuda.empty_cache()              # the eager warm-up's cached blocks would sit beside the graph's private pool
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph, stream=side):
                        out = model(self.left, self.right)
                torch.cuda.current_stream(dev).wait_stream(side)
            self.out: Maps = tuple(out)
        self.copy_stream = torch.cuda.Stream(dev)
        self._stage_in = [(torch.empty_like(self.left), torch.empty_like(self.right)) for _ in range(2)]
        self._stage_out = [[torch.empty_like(o) for o in self.out] for _ in range(2)]
        self._host_out = [[torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in self.out] for _ in range(2)]
        mk = lambda: [torch.cuda.Event() for _ in range(2)]
        self._e_h2d, self._e_in_free, self._e_out_ready, self._e_out_free = mk(), mk(), mk(), mk()

    # -- device-resident step (inputs already in self.left / self.right)
    def replay(self) -> Maps:
        if self.graph is not None:
            self.graph.replay()
        else:
            with torch.inference_mode():
                self.out = tuple(self.model(self.left, self.right))
        return self.out

    def _check(self, t: torch.Tensor):
        if tuple(t.shape) != self.shape or t.dtype != self.left.dtype or t.is_cuda:
            raise ValueError(f"expected a host tensor {self.shape} {self.left.dtype}, got {tuple(t.shape)} {t.dtype} cuda={t.is_cuda}")

    def _h2d(self, i: int, left_h: torch.Tensor, right_h: torch.Tensor):
        b = i & 1
        cs = self.copy_stream
        with torch.cuda.stream(cs):
            if i >= 2:
                cs.wait_event(self._e_in_free[b])     # the compute stream has consumed this staging slot (batch i-2)
            self._stage_in[b][0].copy_(left_h, non_blocking=True)
            self._stage_in[b][1].copy_(right_h, non_blocking=True)
            self._e_h2d[b].record(cs)

    def run(self, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]], host_out: Optional[Sequence[Sequence[torch.Tensor]]] = None):
        """Enqueue the whole sequence (host→device, forward, device→host for every batch) and return WITHOUT synchronising: the
        outputs are valid once ``copy_stream`` has drained (``runner.copy_stream.synchronize()`` or a stream wait).
        ``host_out[i]``: three pinned host tensors for batch i; the default (the runner's two rotating pinned sets) is only
        allowed for at most two batches."""
        n = len(batches)
        if n == 0:
            return
        if host_out is None and n > 2:
            raise ValueError("run(): with the default host_out only the last two batches would survive (two rotating pinned sets); pass "
                             "one pinned output triple per batch, or use stream() / infer()")
        if host_out is not None and len(host_out) != n:
            raise ValueError(f"run(): {n} batches but {len(host_out)} output triples")
        for l, r in batches:
            self._check(l), self._check(r)
        comp = torch.cuda.current_stream(self.device)
        cs = self.copy_stream
        self._h2d(0, *batches[0])
        for i in range(n):
            b = i & 1
            if i + 1 < n:
                self._h2d(i + 1, *batches[i + 1])                 # overlaps this batch's compute
            comp.wait_event(self._e_h2d[b])
            self.left.copy_(self._stage_in[b][0], non_blocking=True)
            self.right.copy_(self._stage_in[b][1], non_blocking=True)
            self._e_in_free[b].record(comp)
            self.replay()
            if i >= 2:
                comp.wait_event(self._e_out_free[b])              # batch i-2 has left this output staging slot
            for so, o in zip(self._stage_out[b], self.out):
                so.copy_(o, non_blocking=True)
            self._e_out_ready[b].record(comp)
            dst = host_out[i] if host_out is not None else self._host_out[b]
            with torch.cuda.stream(cs):
                cs.wait_event(self._e_out_ready[b])
                for h, so in zip(dst, self._stage_out[b]):
                    h.copy_(so, non_blocking=True)                # overlaps the next batch's compute
                self._e_out_free[b].record(cs)

    def stream(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> Iterator[List[torch.Tensor]]:
        """Generator form: yields the three pinned host maps of every batch, in order, one batch behind the submission."""
        prev = None
        comp = torch.cuda.current_stream(self.device)
        i = 0
        it = iter(batches)
        nxt = next(it, None)
        if nxt is not None:
            self._check(nxt[0]), self._check(nxt[1])
            self._h2d(0, *nxt)
        while nxt is not None:
            cur, b = nxt, i & 1
            nxt = next(it, None)
            if nxt is not None:
                self._check(nxt[0]), self._check(nxt[1])
                self._h2d(i + 1, *nxt)
            comp.wait_event(self._e_h2d[b])
            self.left.copy_(self._stage_in[b][0], non_blocking=True)
            self.right.copy_(self._stage_in[b][1], non_blocking=True)
            self._e_in_free[b].record(comp)
            self.replay()
            if i >= 2:
                comp.wait_event(self._e_out_free[b])
            for so, o in zip(self._stage_out[b], self.out):
                so.copy_(o, non_blocking=True)
            self._e_out_ready[b].record(comp)
            if prev is not None:                                   # hand out batch i-1 while batch i computes
                self._e_out_free[prev].synchronize()
                yield self._host_out[prev]
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self._e_out_ready[b])
                for h, so in zip(self._host_out[b], self._stage_out[b]):
                    h.copy_(so, non_blocking=True)
                self._e_out_free[b].record(self.copy_stream)
            prev = b
            i += 1
        if prev is not None:
            self._e_out_free[prev].synchronize()
            yield self._host_out[prev]

    def infer(self, left_h: torch.Tensor, right_h: torch.Tensor) -> List[torch.Tensor]:
        """One synchronous batch: pinned or pageable host images in, three host maps out."""
        self.run([(left_h, right_h)])
        self._e_out_free[0].synchronize()
        return [t.clone() for t in self._host_out[0]]
