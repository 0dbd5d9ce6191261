"""Sparse (x) sparse elementwise operators (reference python/dgl/sparse/elementwise_op_sp.py,
tests/python/pytorch/sparse/test_elementwise_op_sp.py)."""
import pytest
import torch

import dgl_b200 as dgl
import dgl_b200.sparse as dglsp


def _pair():
    A = dglsp.from_coo(torch.tensor([1, 0, 2]), torch.tensor([0, 3, 2]), torch.tensor([1.0, 2.0, 3.0]), (3, 4))
    B = dglsp.from_coo(torch.tensor([0, 2, 2]), torch.tensor([3, 2, 0]), torch.tensor([10.0, 20.0, 30.0]), (3, 4))
    return A, B


def test_mul_is_the_intersection_product():
    A, B = _pair()
    C = A * B
    assert torch.allclose(C.to_dense(), A.to_dense() * B.to_dense()) and C.nnz == 2
    assert torch.allclose(dglsp.sp_mul(A, B).to_dense(), C.to_dense())
    assert (A * 2).val.tolist() == [2.0, 4.0, 6.0] and (2 * A).val.tolist() == [2.0, 4.0, 6.0]
    assert dglsp.sp_mul(A, torch.tensor(3.0)).val.tolist() == [3.0, 6.0, 9.0]
    empty = dglsp.from_coo(torch.tensor([0]), torch.tensor([0]), torch.tensor([1.0]), (3, 4))
    assert (A * empty).nnz == 0
    with pytest.raises(dgl.DGLError):
        A * dglsp.identity((3, 3))
    eye = dglsp.identity((3, 3))
    assert (eye * eye).val.tolist() == [1.0, 1.0, 1.0]


def test_mul_gradients_flow_to_both_operands():
    A, B = _pair()
    A.val.requires_grad_(True)
    B.val.requires_grad_(True)
    (A * B).val.sum().backward()
    assert A.val.grad.tolist() == [0.0, 10.0, 20.0] and B.val.grad.tolist() == [2.0, 3.0, 0.0]


def test_div_requires_identical_sparsity_and_keeps_order():
    A, B = _pair()
    D = dglsp.val_like(A, torch.tensor([2.0, 4.0, 6.0]))
    assert (D / A).val.tolist() == [2.0, 2.0, 2.0]
    # same pattern, different storage order
    P = dglsp.from_coo(torch.tensor([2, 1, 0]), torch.tensor([2, 0, 3]), torch.tensor([3.0, 1.0, 2.0]), (3, 4))
    q = A / P
    assert q.val.tolist() == [1.0, 1.0, 1.0] and torch.equal(q.row, A.row)
    with pytest.raises(dgl.DGLError, match="different sparsities"):
        A / B
    assert (A / 2).val.tolist() == [0.5, 1.0, 1.5]
    with pytest.raises(dgl.DGLError):
        dglsp.div(2.0, A)


def test_operator_forms_and_scalars():
    A, B = _pair()
    assert dglsp.sp_add(A, 1) is NotImplemented and dglsp.sp_sub(A, 1.0) is NotImplemented
    assert dglsp.sp_mul(A, "x") is NotImplemented and dglsp.sp_power(A, A) is NotImplemented
    assert torch.allclose(dglsp.sp_add(A, B).to_dense(), A.to_dense() + B.to_dense())
    assert torch.allclose(dglsp.sp_sub(A, B).to_dense(), A.to_dense() - B.to_dense())
    assert torch.allclose(dglsp.spsp_add(A, B).to_dense(), A.to_dense() + B.to_dense())
    assert dglsp.sp_power(A, 2).val.tolist() == [1.0, 4.0, 9.0]
    assert dglsp.is_scalar(3) and dglsp.is_scalar(torch.tensor(1.0)) and not dglsp.is_scalar(torch.ones(1))
    dglsp.load_dgl_sparse()


def test_bsddmm_broadcasts_scalar_values_over_the_batch():
    A = dglsp.from_coo(torch.tensor([0, 1, 2]), torch.tensor([1, 2, 0]), torch.tensor([1.0, 2.0, 3.0]), (3, 3))
    X1, X2 = torch.randn(3, 4, 5), torch.randn(4, 3, 5)
    out = dglsp.bsddmm(A, X1, X2)
    dense = torch.einsum("mkb,knb->mnb", X1, X2)
    r, c = A.coo()
    assert out.val.shape == (3, 5) and torch.allclose(out.val, A.val[:, None] * dense[r.long(), c.long()], atol=1e-5)
    B = dglsp.val_like(A, torch.rand(3, 5))
    out2 = dglsp.bsddmm(B, X1, X2)
    assert torch.allclose(out2.val, B.val * dense[r.long(), c.long()], atol=1e-5)
