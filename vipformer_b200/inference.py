"""Frozen-feature extraction: the evaluation half of the reference (pretrain.py:228-276, eval_fewshot.py, eval_zeroshot.py).

    for data, label in loader:  feats = pc_model(data)[1]      # eval mode, N = 1024 points, backbone features [B, 2D]
    ... sklearn SVC on the concatenated features

Here: eval mode + no_grad (no context is saved, dropout off), eval-mode BatchNorm folded into the preceding convolution
(functional.group2emb_fwd), fixed-size batches replayed from ONE CUDA graph, features written straight into a
device-resident result matrix -- no `.tolist()` / per-batch host round trip (pretrain.py:243-246); one device-to-host
copy at the end (pinned) for the CPU-side classifier.
"""
import torch

from . import _lib

F32 = torch.float32


class FeatureExtractor:
    def __init__(self, model, batch_size, num_points, *, use_cuda_graph=True, seed=1234, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.model = model.to(self.device).eval()
        self.B, self.N = int(batch_size), int(num_points)
        self.pts = torch.zeros((self.B, self.N, 3), dtype=F32, device=self.device)
        self.start = torch.zeros(self.B, dtype=torch.long, device=self.device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(seed)
        self.use_graph, self.graph, self.out = use_cuda_graph, None, None
        self.fixed_start = None      # set to a LongTensor [B] to pin the FPS start points (tests)

    def _body(self):
        if self.fixed_start is None:    # utils.py:71: the reference draws the FPS start from torch's RNG also at eval time
            torch.randint(0, self.N, (self.B,), dtype=torch.long, device=self.device, generator=self.gen, out=self.start)
        else:
            self.start.copy_(self.fixed_start)
        self.model.fps_start_idx = self.start
        with torch.no_grad():
            _, backbone = self.model(self.pts)
        return backbone

    def _run(self):
        if not self.use_graph:
            return self._body()
        if self.graph is None:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream())
            gstate = self.gen.get_state()
            with torch.cuda.stream(s):
                self._body()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.gen.set_state(gstate)
            self.graph = torch.cuda.CUDAGraph()
            self.graph.register_generator_state(self.gen)
            with torch.cuda.graph(self.graph):
                self.out = self._body()
        self.graph.replay()
        return self.out

    @torch.no_grad()
    def __call__(self, points):
        """points [n, N, 3] fp32 (host, preferably pinned, or device) -> backbone features [n, 2D] fp32 on the device."""
        n = points.shape[0]
        if points.shape[1] != self.N:
            raise ValueError(f"extractor was built for {self.N} points per cloud, got {points.shape[1]}")
        feats = None
        for i0 in range(0, n, self.B):
            m = min(self.B, n - i0)
            self.pts[:m].copy_(points[i0:i0 + m], non_blocking=True)
            if m < self.B:                       # ragged tail: repeat the last cloud (BatchNorm is in eval mode: rows independent)
                self.pts[m:].copy_(self.pts[m - 1:m].expand(self.B - m, -1, -1))
            out = self._run()
            if feats is None:
                feats = torch.empty((n, out.shape[1]), dtype=F32, device=self.device)
            feats[i0:i0 + m].copy_(out[:m])
        return feats

    @torch.no_grad()
    def extract(self, loader):
        """loader yields (points [b, N, 3], label [b, ...]) like ModelNet40SVM / ScanObjectNNSVM (datasets/data.py:120-146).
        -> (features [n, 2D], labels [n]) as numpy arrays for sklearn (pretrain.py:247-251)."""
        fs, ls = [], []
        for data, label in loader:
            fs.append(self(data))
            ls.append(torch.as_tensor(label).reshape(-1))
        f = torch.cat(fs)
        host = torch.empty(f.shape, dtype=F32).pin_memory()
        host.copy_(f, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host.numpy(), torch.cat(ls).numpy()
