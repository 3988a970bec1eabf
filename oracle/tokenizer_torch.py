"""Device-agnostic plain-PyTorch restatement of the reference tokenizer -- TEST / BENCH INFRASTRUCTURE ONLY.

What the reference executes on a GPU today: the same ATen ops, in the same order, as
vipformer/model/pointcloud/utils.py (farthest_point_sample:56-85 -- a Python loop of npoint iterations of gather /
sub / square / sum / min / argmax; square_distance:122-141 -- expanded form with a K = 3 bmm; knn_point:107-119 --
torch.topk(largest=False); divide_patches:6-38 -- flat gather and the slot-axis centre subtraction of :36).
bench.py's `gpu_eager_baseline` leg times it (with oracle/model_ref.py under bf16 autocast) on the same B200 as the
stronger comparator SURVEY.md 8(d) asks for; tests/test_oracle_golden.py pins it to the C oracle under the stated
(distance, index) tie rule.
"""
import torch


def farthest_point_sample(pts, npoint, start_idx):
    B, N, _ = pts.shape
    dev = pts.device
    centroids = torch.zeros(B, npoint, dtype=torch.long, device=dev)
    distance = torch.full((B, N), 1e10, device=dev, dtype=pts.dtype)
    farthest = start_idx.to(dev).long()
    batch = torch.arange(B, dtype=torch.long, device=dev)
    xyz = pts[:, :, :3]
    for i in range(npoint):
        centroids[:, i] = farthest
        c = xyz[batch, farthest, :].view(B, 1, 3)
        dist = ((xyz - c) ** 2).sum(-1)
        distance = torch.minimum(distance, dist)
        farthest = distance.max(-1)[1]
    return centroids


def square_distance(src, dst):
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist = dist + (src ** 2).sum(-1)[:, :, None]
    dist = dist + (dst ** 2).sum(-1)[:, None, :]
    return dist


def divide_patches(points, num_groups, group_size, start_idx, sorted_knn=True):
    B, N, C = points.shape
    idx = farthest_point_sample(points, num_groups, start_idx)
    centers = torch.gather(points, 1, idx[:, :, None].expand(-1, -1, C))
    with torch.autocast(points.device.type, enabled=False):        # the tokenizer stays fp32
        d = square_distance(centers[:, :, :3].float(), points[:, :, :3].float())
    nidx = torch.topk(d, group_size, dim=-1, largest=False, sorted=sorted_knn)[1]
    flat = (nidx + torch.arange(B, device=points.device).view(-1, 1, 1) * N).view(-1)
    nb = points.reshape(B * N, C)[flat].reshape(B, num_groups, group_size, C).clone()
    nb[:, :, :3] = nb[:, :, :3] - centers.unsqueeze(2)[:, :, :3]      # slot-axis slice: utils.py:36
    return nb, centers
