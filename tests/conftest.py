import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    try:
        import torch
        torch.set_num_threads(1)   # torch CPU sorts crawl when their threads oversubscribe a small cgroup
    except Exception:
        pass


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the product library, the C oracle and (where /root/reference exists) oracle/_ref are built."""
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(ROOT, "col_bwt_b200", "libcolbwt_b200.so")) or not os.path.exists(
            os.path.join(ROOT, "oracle", "liboracle.so")):
        g.build()
    yield


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def small_index(tmp_path_factory):
    """4 haplotypes x 30 kbp (+revcomp), tunnels marks every 5 columns, with reads (1 % substitutions)."""
    from synthdata import pangenome as P, pipeline as PL
    haps = P.make_haplotypes(30000, 4, snp=2e-3, indel=2e-4, seed=11)
    idx = PL.build_index(haps, split_rate=5)
    d = tmp_path_factory.mktemp("small")
    path = str(d / "small.fa.col_pml")
    PL.write_col_pml(path, idx["columns"])
    seqs, off = P.sample_reads(idx["text"], idx["seq_starts"], 3000, 150, sub=0.01, seed=3)
    return {"path": path, "cols": idx["columns"], "idx": idx, "seqs": seqs, "off": off, "haps": haps}
