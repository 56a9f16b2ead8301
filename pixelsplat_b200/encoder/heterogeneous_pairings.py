""""All other views" index tables (/root/reference/src/misc/heterogeneous_pairings.py:9-43)."""
import torch


def generate_heterogeneous_index(n: int, device=torch.device("cpu")):
    """index_self[h, w] = h;  index_other[h, w] = w if w < h else w + 1   (shape [n, n-1])."""
    a = torch.arange(n, device=device)
    index_self = a[:, None].expand(n, n - 1).clone()
    w = a[None, : n - 1].expand(n, n - 1)
    index_other = w + (w >= a[:, None]).long()
    return index_self, index_other


def generate_heterogeneous_index_transpose(n: int, device=torch.device("cpu")):
    """Indices (t_v, t_ov) that swap the roles of view and other-view; an involution."""
    a = torch.arange(n, device=device)
    w = a[None, : n - 1].expand(n, n - 1)
    h = a[:, None].expand(n, n - 1)
    t_v = w + (w >= h).long()          # the other view itself
    t_ov = h - (w < h).long()          # position of h among the others of that view
    return t_v, t_ov
