"""Host sparse-structure ops (csrc/cpu/spmat.cc) against torch oracles, for both id widths."""
import pytest
import torch

import dgl_b200 as dgl
from dgl_b200.ops import cpu_struct


@pytest.fixture(params=[torch.int32, torch.int64])
def coo(request):
    torch.manual_seed(7)
    idt, n, E = request.param, 700, 15000
    return idt, n, torch.randint(0, n, (E,), dtype=idt), torch.randint(0, n, (E,), dtype=idt)


def test_coo_to_csr_is_a_stable_sort(coo):
    idt, n, row, col = coo
    indptr, indices, eids = cpu_struct.coo_to_csr(row, col, n)
    perm = torch.sort(row, stable=True)[1]
    assert indptr.dtype == idt and torch.equal(eids.long(), perm) and torch.equal(indices, col[perm])
    want = torch.zeros(n + 1, dtype=torch.int64)
    want[1:] = torch.cumsum(torch.bincount(row.long(), minlength=n), 0)
    assert torch.equal(indptr.long(), want)
    e_indptr, e_idx, e_eid = cpu_struct.coo_to_csr(row[:0], col[:0], 5)
    assert e_indptr.tolist() == [0] * 6 and e_idx.numel() == 0 and e_eid.numel() == 0


def test_transpose_round_trip(coo):
    idt, n, row, col = coo
    indptr, indices, eids = cpu_struct.coo_to_csr(row, col, n)
    t_indptr, t_indices, t_data = cpu_struct.csr_transpose(indptr, indices, eids, n)
    c_indptr, _, _ = cpu_struct.coo_to_csr(col, row, n)
    assert torch.equal(t_indptr, c_indptr)
    assert torch.equal(row[t_data.long()], t_indices)
    dst_of_slot = torch.repeat_interleave(torch.arange(n), (t_indptr[1:] - t_indptr[:-1]).long())
    assert torch.equal(col[t_data.long()].long(), dst_of_slot)
    # inside a column, rows ascend (the stable sort keeps CSR order)
    same = dst_of_slot[1:] == dst_of_slot[:-1]
    assert bool((t_indices[1:][same] >= t_indices[:-1][same]).all())
    tp, ti, perm = cpu_struct.transpose(indptr, indices, n)
    assert torch.equal(tp, t_indptr) and torch.equal(ti, t_indices) and perm.dtype == torch.int64


def test_coalesce_matches_unique(coo):
    idt, n, row, col = coo
    r, c, group, count, first = cpu_struct.coalesce(row, col, n, n)
    key = row.long() * n + col.long()
    uniq, inverse, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    assert torch.equal(r.long() * n + c.long(), uniq) and torch.equal(group, inverse) and torch.equal(count, cnt)
    f = torch.full((uniq.shape[0],), key.shape[0])
    f.scatter_reduce_(0, inverse, torch.arange(key.shape[0]), reduce="amin")
    assert torch.equal(first, f)


def test_slices(coo):
    idt, n, row, col = coo
    indptr, indices, eids = cpu_struct.coo_to_csr(row, col, n)
    rows = torch.tensor([5, 3, n - 1, 5], dtype=idt)
    sp, si, sd = cpu_struct.slice_rows(indptr, indices, eids, rows)
    for i, r in enumerate(rows.tolist()):
        assert torch.equal(si[sp[i]:sp[i + 1]], indices[indptr[r]:indptr[r + 1]])
        assert torch.equal(sd[sp[i]:sp[i + 1]], eids[indptr[r]:indptr[r + 1]])
    keep = torch.randperm(n)[:200]
    col_map = torch.full((n,), -1, dtype=idt)
    col_map[keep] = torch.arange(200, dtype=idt)
    sp, si, sd = cpu_struct.slice_matrix(indptr, indices, eids, keep.to(idt), col_map)
    m = (col_map[row.long()] >= 0) & (col_map[col.long()] >= 0)
    assert torch.equal(torch.sort(sd.long())[0], torch.nonzero(m).squeeze(1))
    assert torch.equal(si, col_map[col[sd.long()].long()])
    src_new = torch.repeat_interleave(torch.arange(200), (sp[1:] - sp[:-1]).long())
    assert torch.equal(col_map[row[sd.long()].long()].long(), src_new)


def test_graph_level_users_agree_with_the_tensor_paths():
    torch.manual_seed(2)
    g = dgl.rand_graph(300, 4000)
    # node_subgraph: native slice (CSR present, small node set) == masked scan
    nodes = torch.randperm(300)[:40]
    a = dgl.node_subgraph(g, nodes)
    g.create_formats_()
    b = dgl.node_subgraph(g, nodes)
    assert torch.equal(a.edata[dgl.EID], b.edata[dgl.EID])
    for x, y in zip(a.edges(), b.edges()):
        assert torch.equal(x, y)
    # to_simple
    sg, mapping = dgl.to_simple(g, return_counts="cnt", writeback_mapping=True)
    u, v = g.edges()
    key = u * 300 + v
    uniq, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    su, sv = sg.edges()
    assert torch.equal(su * 300 + sv, uniq) and torch.equal(mapping, inv) and torch.equal(sg.edata["cnt"], cnt)
    # line graph, with and without backtracking
    small = dgl.graph((torch.tensor([0, 1, 1, 2, 0]), torch.tensor([1, 0, 2, 0, 2])))
    for bt in (True, False):
        lg = dgl.line_graph(small, backtracking=bt)
        got = set(zip(*[t.tolist() for t in lg.edges()]))
        su, sv = [t.tolist() for t in small.edges()]
        want = {(i, j) for i in range(5) for j in range(5) if sv[i] == su[j] and (bt or not (sv[j] == su[i]))}
        assert got == want
