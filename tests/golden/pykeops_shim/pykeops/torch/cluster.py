"""pykeops.torch.cluster — dense restatement of the block-sparse helpers (public API semantics)."""
import torch


def grid_cluster(x, size):
    """Voxel labels of the points x:(N,D), D <= 3: ``floor((x - min) / size)`` per axis, mixed with the weights
    (2^20, 2^10, 1) and relabelled to a compact 0..C-1 range in increasing order of the mixed key; int32."""
    with torch.no_grad():
        D = x.shape[1]
        if D == 1:
            weights = torch.IntTensor([1])
        elif D == 2:
            weights = torch.IntTensor([2**10, 1])
        elif D == 3:
            weights = torch.IntTensor([2**20, 2**10, 1])
        else:
            raise NotImplementedError()
        x_ = ((x - x.min(0)[0]) / size).floor().int()
        x_ = x_ * weights.to(x.device)
        lab = x_.sum(1)
        lab = lab - lab.min()
        u_lab = torch.unique(lab).sort()[0]
        n_lab = len(u_lab)
        foo = torch.empty(int(u_lab.max()) + 1, dtype=torch.int32, device=x.device)
        foo[u_lab.long()] = torch.arange(n_lab, dtype=torch.int32, device=x.device)
        lab = foo[lab.long()]
    return lab


def cluster_ranges(lab, Nlab=None):
    """[start, end) index ranges of the clusters once the points are SORTED by label; int32 (C,2)."""
    if Nlab is None:
        Nlab = torch.bincount(lab.long()).float()
    pivots = torch.cat((torch.zeros(1, device=Nlab.device), Nlab.cumsum(0)))
    return torch.stack((pivots[:-1], pivots[1:])).t().int()


def cluster_centroids(x, lab, Nlab=None, weights=None, weights_c=None):
    if Nlab is None:
        Nlab = torch.bincount(lab.long()).float()
    if weights is not None and weights_c is None:
        weights_c = torch.bincount(lab.long(), weights=weights).view(-1, 1)
    c = torch.zeros((len(Nlab), x.shape[1]), dtype=x.dtype, device=x.device)
    for d in range(x.shape[1]):
        if weights is None:
            c[:, d] = torch.bincount(lab.long(), weights=x[:, d]) / Nlab
        else:
            c[:, d] = torch.bincount(lab.long(), weights=x[:, d] * weights.view(-1)) / weights_c.view(-1)
    return c


def cluster_ranges_centroids(x, lab, weights=None, min_weight=1e-9):
    """(ranges (C,2) int32, centroids (C,D) weight-averaged, weights_c (C,) summed)."""
    Nlab = torch.bincount(lab.long()).float()
    if weights is not None:
        w_c = torch.bincount(lab.long(), weights=weights).view(-1)
        w_c[w_c.abs() <= min_weight] = min_weight
    else:
        w_c = None
    ranges = cluster_ranges(lab, Nlab)
    x_c = cluster_centroids(x, lab, Nlab, weights=weights, weights_c=w_c)
    if weights is None:
        return ranges, x_c
    return ranges, x_c, w_c


def sort_clusters(x, lab):
    lab, perm = torch.sort(lab.view(-1))
    if type(x) is tuple:
        x_sorted = tuple(a[perm] for a in x)
    elif type(x) is list:
        x_sorted = list(a[perm] for a in x)
    else:
        x_sorted = x[perm]
    return x_sorted, lab


def from_matrix(ranges_i, ranges_j, keep):
    """Cluster-level boolean matrix -> the 6-tuple (ranges_i, slices_i, redranges_j, ranges_j, slices_j,
    redranges_i) of a block-sparse reduction."""
    I, J = torch.meshgrid(torch.arange(0, keep.shape[0]), torch.arange(0, keep.shape[1]), indexing="ij")
    redranges_i = ranges_i[I.t()[keep.t()]]
    redranges_j = ranges_j[J[keep]]
    slices_i = keep.sum(1).cumsum(0).int()
    slices_j = keep.sum(0).cumsum(0).int()
    return (ranges_i, slices_i, redranges_j, ranges_j, slices_j, redranges_i)


def swap_axes(ranges):
    return (*ranges[3:6], *ranges[0:3])


def ranges_to_mask(ranges, n_i, n_j, axis=1):
    """Point-level boolean mask (n_i, n_j) of the pairs a block-sparse reduction over ``axis`` visits."""
    if axis == 0:
        return ranges_to_mask(swap_axes(ranges), n_j, n_i, axis=1).t()
    ranges_i, slices_i, redranges_j = ranges[0], ranges[1], ranges[2]
    mask = torch.zeros(n_i, n_j, dtype=torch.bool)
    start = 0
    for k in range(len(ranges_i)):
        i0, i1 = int(ranges_i[k, 0]), int(ranges_i[k, 1])
        end = int(slices_i[k])
        if end > start:
            cols = torch.zeros(n_j, dtype=torch.bool)
            for j0, j1 in redranges_j[start:end].tolist():
                cols[j0:j1] = True
            mask[i0:i1] = cols
        start = end
    return mask
