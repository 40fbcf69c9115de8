"""The tests of tests/test_xpending_gpu.py for the metrics added last (written without a GPU) executed on the CPU with the oracle-backed engine in
place of the CUDA one: the test CODE -- fixtures, KAT keys, tolerances against the reference's matrices -- is exercised before its
first hardware run.  (On the GPU the same bodies compare the CUDA kernels with the oracle bit for bit.)"""
import numpy as np
import pytest


@pytest.fixture
def oracle_backed(monkeypatch):
    import pynndescent_b200.engine as eng_mod
    from oracle_engine import OracleEngine
    import pynndescent_b200.pynndescent_ as nnd_mod
    monkeypatch.setattr(eng_mod, "Engine", OracleEngine)
    monkeypatch.setattr(nnd_mod, "Engine", OracleEngine)

    class _NoCuda:                                   # stands in for the torch_cuda fixture (torch_cuda.cuda.synchronize())
        class cuda:
            @staticmethod
            def synchronize():
                pass
    return _NoCuda


@pytest.mark.parametrize("p", [0.5, 1.5, 3.0, 4.0])
def test_minkowski_pair_distance_body(oracle_backed, oracle, p):
    import test_xpending_gpu as tx
    tx.test_minkowski_pair_distances(oracle_backed, oracle, p)


def test_inner_product_true_angular_kat_body(oracle_backed):
    import test_xpending_gpu as tx
    tx.test_inner_product_and_true_angular_kat_vs_reference(oracle_backed)


@pytest.mark.parametrize("metric", ["inner_product", "true_angular"])
def test_pair_distance_body(oracle_backed, oracle, metric):
    import test_xpending_gpu as tx
    tx.test_pair_distances_bit_exact_vs_oracle(oracle_backed, oracle, metric, 37)


@pytest.mark.parametrize("name", ["latent_inner_1500x20", "latent_trueang_1500x20", "latent_minkowski3_1500x20"])
def test_staged_build_bodies(oracle_backed, oracle, name):
    """leaf init from the reference's leaves, init_random + two candidate / join iterations, the whole build: the staged GPU tests'
    bodies, golden keys and metric_kwds threading, with every stage executed by the oracle."""
    import test_xpending_gpu as tx
    tx.test_staged_builds_from_reference_goldens(oracle_backed, oracle, name)


@pytest.mark.parametrize("metric,n,d,k,seed", [("inner_product", 600, 20, 8, 13), ("true_angular", 600, 24, 8, 14)])
def test_full_build_body(oracle_backed, oracle, metric, n, d, k, seed):
    import test_xpending_gpu as tx
    tx.test_full_build_bit_exact_vs_oracle(oracle_backed, oracle, metric, n, d, k, seed)
