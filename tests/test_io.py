"""MatrixUtils::Dump (MatrixMarket) and Preconditioner::Visualize (MATLAB file with the partitioning), SURVEY 8f-4."""
import re

import numpy as np
import scipy.io
import scipy.sparse as sp

import hymls_b200 as hb
from oracle.partitioner import OverlappingPartitioner
from tests.common import make_params
from tests.test_host_maps import _dictify


def test_dump_round_trips_through_matrixmarket(tmp_path):
    A = sp.csr_matrix(hb.galeri.create_matrix("Stokes-C", 2, 8))
    hb.io.dump_matrix(A, str(tmp_path / "A.mtx"))
    B = scipy.io.mmread(str(tmp_path / "A.mtx")).tocsr()
    assert (abs(A - B)).max() == 0.0 and B.nnz == A.nnz
    X = np.random.default_rng(0).uniform(-1, 1, (A.shape[0], 2))
    hb.io.dump_vector(X, str(tmp_path / "x.mtx"))
    assert np.array_equal(scipy.io.mmread(str(tmp_path / "x.mtx")), X)
    with open(tmp_path / "A.mtx") as f:
        assert f.readline().strip() == "%%MatrixMarket matrix coordinate real general"


def test_visualize_writes_the_groups_of_every_level(tmp_path):
    p = make_params("Stokes-C", 2, 16, 4, 2, 2)
    A = hb.galeri.create_matrix("Stokes-C", 2, 16)
    P = hb.Preconditioner(A, _dictify(p), pattern_only=True)
    P.Initialize()
    out = str(tmp_path / "hid_data.m")
    P.Visualize(out)
    text = open(out).read()
    assert "% Domain decomposition and separators, level 0" in text and "level 1" in text
    hid = OverlappingPartitioner(p.copy(), 0)
    # subdomain 3 of level 0: interior and every separator group, in the reference's order and syntax
    m = re.search(r"p\{0\}\{1\}\.groups\{4\} = \{(.*?)\};", text, re.S)
    lists = [[int(t) for t in g.split(",") if t] for g in re.findall(r"\[([^\]]*)\]", m.group(1))]
    assert lists[0] == list(hid.interior[3])
    assert lists[1:] == [list(n) for _, n in hid.groups[3]]
    vs = re.search(r"p\{0\}\{1\}\.vsums=\[(.*?)\];", text).group(1).split()
    assert [int(v) for v in vs] == [int(v) for v in P.GetMap(hb.api.MAP_VSUM, 0)]
