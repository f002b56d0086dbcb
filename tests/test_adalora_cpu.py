"""Host arithmetic of the AdaLoRA adapter (neuspeech1_b200/adalora.py) that needs no GPU: the orthogonality regulariser's
closed-form gradient against autograd."""
import torch

from neuspeech1_b200.adalora import orth_regulariser


def test_orth_regulariser_gradient_matches_autograd():
    g = torch.Generator().manual_seed(0)
    A = (torch.randn(5, 12, 64, generator=g) * 0.1).requires_grad_(True)
    B = (torch.randn(5, 96, 12, generator=g) * 0.1).requires_grad_(True)
    eye = torch.eye(12)
    ref = sum(torch.norm(A[i] @ A[i].T - eye, p="fro") + torch.norm(B[i].T @ B[i] - eye, p="fro") for i in range(5))
    ref.backward()
    s, dA, dB = orth_regulariser(A.detach(), B.detach())
    assert torch.allclose(s, ref.detach(), rtol=1e-6)
    assert torch.allclose(dA, A.grad, rtol=1e-5, atol=1e-7) and torch.allclose(dB, B.grad, rtol=1e-5, atol=1e-7)
