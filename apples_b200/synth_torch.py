"""Large-scale synthetic inputs generated on the GPU with torch (bench.py; SURVEY.md section 8d shapes).

Same model as apples_b200.synth (JC69 evolution down the tree, 3 % point gaps, leading/trailing gap runs of
U(0, 0.1 L), queries = a random leaf evolved by a further Exp(mean 0.05)) but level-batched on the device so that the
200 000-leaf x 5000-site configuration takes seconds.  torch is plumbing here (device RNG and memory), not the product.
"""
import numpy as np
import torch

_GAP = 45  # '-'


def _sub_prob(t, ns=4):
    a = ns / (ns - 1.0)
    return (1.0 / a) * (1.0 - torch.exp(-a * torch.clamp(t, min=0.0)))


def _mutate_rows(states, t, gen, chunk=16384, ns=4):
    """states uint8 [n, L] (device), t float32 [n] branch lengths -> mutated copy (ns-state symmetric model)"""
    n, L = states.shape
    out = torch.empty_like(states)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        p = _sub_prob(t[a:b], ns).unsqueeze(1)
        hit = torch.rand((b - a, L), device=states.device, generator=gen) < p
        shift = torch.randint(1, ns, (b - a, L), device=states.device, generator=gen, dtype=torch.uint8)
        s = states[a:b]
        out[a:b] = torch.where(hit, (s + shift) & 3 if ns == 4 else (s + shift) % ns, s)
    return out


def _gapify_rows(states, gen, gap_frac, edge_frac, chunk=16384, alphabet=b'ACGT'):
    """states uint8 codes [n, L] -> ASCII bytes with gaps"""
    n, L = states.shape
    alpha = torch.tensor(list(alphabet), dtype=torch.uint8, device=states.device)
    out = torch.empty_like(states)
    col = torch.arange(L, device=states.device).unsqueeze(0)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        ch = alpha[states[a:b].long()]
        g = torch.rand((b - a, L), device=states.device, generator=gen) < gap_frac
        lead = torch.randint(0, int(edge_frac * L) + 1, (b - a, 1), device=states.device, generator=gen)
        trail = torch.randint(0, int(edge_frac * L) + 1, (b - a, 1), device=states.device, generator=gen)
        g = g | (col < lead) | (col >= L - trail)
        out[a:b] = torch.where(g, torch.full_like(ch, _GAP), ch)
    return out


AA_ALPHABET = b'ARNDCQEGHILKMFPSTWYV'


def evolve_alignment(tree, L, seed, device, gap_frac=0.03, edge_frac=0.1, protein=False):
    """Returns (ref_bytes uint8 [n_leaves, L] on `device`, rows ordered by leaf node id; leaf_states uint8 [n_leaves, L])."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    ns = 20 if protein else 4
    alphabet = AA_ALPHABET if protein else b'ACGT'
    M = tree.num_nodes
    S = torch.empty((M, L), dtype=torch.uint8, device=device)
    S[M - 1] = torch.randint(0, ns, (L,), device=device, generator=gen, dtype=torch.uint8)
    level = tree.level
    order = np.argsort(level, kind='stable')
    lv_sorted = level[order]
    bounds = np.searchsorted(lv_sorted, np.arange(1, level.max() + 2))
    par = torch.from_numpy(tree.parent.astype(np.int64)).to(device)
    el = torch.from_numpy(tree.edge_length.astype(np.float32)).to(device)
    for k in range(len(bounds) - 1):
        idx = torch.from_numpy(order[bounds[k]:bounds[k + 1]].astype(np.int64)).to(device)
        if idx.numel() == 0:
            continue
        S[idx] = _mutate_rows(S[par[idx]], el[idx], gen, ns=ns)
    leaves = torch.from_numpy(tree.leaf_ids.astype(np.int64)).to(device)
    leaf_states = S[leaves].contiguous()
    del S
    return _gapify_rows(leaf_states, gen, gap_frac, edge_frac, alphabet=alphabet), leaf_states


def make_queries(leaf_states, n_queries, seed, device, mean_extra=0.05, gap_frac=0.03, edge_frac=0.1, protein=False):
    """Returns (query_bytes uint8 [n_queries, L] on device, source leaf row index int64 [n_queries])."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_leaves = leaf_states.shape[0]
    src = torch.randint(0, n_leaves, (n_queries,), device=device, generator=gen)
    extra = -mean_extra * torch.log1p(-torch.rand((n_queries,), device=device, generator=gen))
    out = torch.empty((n_queries, leaf_states.shape[1]), dtype=torch.uint8, device=device)
    chunk = 16384
    for a in range(0, n_queries, chunk):
        b = min(n_queries, a + chunk)
        st = _mutate_rows(leaf_states[src[a:b]], extra[a:b].float(), gen, ns=20 if protein else 4)
        out[a:b] = _gapify_rows(st, gen, gap_frac, edge_frac, alphabet=AA_ALPHABET if protein else b'ACGT')
    return out, src


def pack_nucleotide(chars, chunk=8192):
    """ASCII bytes uint8 [n, L] on the device -> packed planes uint32-as-int32 [n, 3, W] on the device
    (same layout as apples_b200.fasta.pack_nucleotide)."""
    n, L = chars.shape
    W = ((L + 31) // 32 + 3) // 4 * 4
    dev = chars.device
    lut = torch.full((256,), 4, dtype=torch.uint8, device=dev)
    for i, c in enumerate(b'ACGT'):
        lut[c] = i
    weights = (torch.ones(32, dtype=torch.int64, device=dev) << torch.arange(32, device=dev)).view(1, 1, 32)
    out = torch.zeros((n, 3, W), dtype=torch.int32, device=dev)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        code = lut[chars[a:b].long()]
        pad = torch.full((b - a, W * 32), 4, dtype=torch.uint8, device=dev)
        pad[:, :L] = code
        valid = pad < 4
        for k, bits in enumerate(((pad & 1).bool() & valid, ((pad >> 1) & 1).bool() & valid, valid)):
            w = (bits.view(b - a, W, 32).long() * weights).sum(dim=2)
            out[a:b, k] = ((w + 2 ** 31) % 2 ** 32 - 2 ** 31).to(torch.int32)
    return out
