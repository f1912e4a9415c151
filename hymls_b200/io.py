"""Debug / inspection output of the reference, on the host side of the C ABI.

  MatrixUtils::Dump(Epetra_CrsMatrix / Epetra_MultiVector, file)   src/HYMLS_MatrixUtils.cpp:559-664
        -> EpetraExt::RowMatrixToMatrixMarketFile / MultiVectorToMatrixMarketFile formats
  Preconditioner::Visualize(mfile) + HierarchicalMap::Print        src/HYMLS_Preconditioner.cpp:753-779,
  + SchurPreconditioner::Visualize                                  src/HYMLS_HierarchicalMap.cpp:339-402,
                                                                    src/HYMLS_SchurPreconditioner.cpp:1624-1652
        -> a MATLAB file with the groups of every subdomain and the V-sum nodes of every level
"""
import numpy as np
import scipy.sparse as sp


def dump_matrix(A, filename, label="HYMLS matrix"):
    """%%MatrixMarket matrix coordinate real general, 1-based, entries row by row (what the reference's
    MatrixUtils::Dump(A, file) writes through EpetraExt::RowMatrixToMatrixMarketFile)"""
    A = sp.csr_matrix(A)
    A.sort_indices()
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    with open(filename, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write("%% %s\n" % label)
        f.write("%d %d %d\n" % (A.shape[0], A.shape[1], A.nnz))
        for r, c, v in zip(rows, A.indices, A.data):
            f.write("%d %d %.16e\n" % (r + 1, c + 1, v))


def dump_vector(x, filename, label="HYMLS vector"):
    """%%MatrixMarket matrix array real general, column major (EpetraExt::MultiVectorToMatrixMarketFile)"""
    X = np.asarray(x, dtype=np.float64)
    X = X.reshape(X.shape[0], -1)
    with open(filename, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n")
        f.write("%% %s\n" % label)
        f.write("%d %d\n" % X.shape)
        for j in range(X.shape[1]):
            for v in X[:, j]:
                f.write("%.16e\n" % v)


def visualize(prec, mfilename, no_recurse=False, rank=0):
    """Preconditioner::Visualize: `prec` is an initialized hymls_b200.Preconditioner (single-rank view)."""
    from . import api
    levels = prec.NumLevels()
    with open(mfilename, "w") as f:
        f.write("\n")
        for lv in range(levels):
            f.write("%" * 54 + "\n")
            f.write("%% Domain decomposition and separators, level %d       %%\n" % lv)
            f.write("%" * 54 + "\n\n")
            f.write("%%Partition %d\n" % rank)
            f.write("%=============\n")
            for sd in range(prec.NumMySubdomains(lv)):
                f.write("p{%d}{%d}.groups{%d} = {" % (lv, rank + 1, sd + 1))
                f.write("[" + "".join("%d," % g for g in prec.GetInteriorGroup(sd, lv)) + "]")
                for _, nodes in prec.GetSeparatorGroups(sd, lv):
                    f.write(",...\n")
                    f.write("[" + "".join("%d," % g for g in nodes) + "]")
                f.write("};\n\n")
            f.write("%" + "$" * 52 + "%\n\n")
            f.write("%% rank %d: SchurPreconditioner (level %d)\n" % (rank, lv))
            f.write("p{%d}{%d}.vsums=[" % (lv, rank + 1))
            f.write("".join("%d " % g for g in prec.GetMap(api.MAP_VSUM, lv)))
            f.write("];\n")
            if no_recurse:
                break
