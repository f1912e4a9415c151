// Drives include/HYMLS_B200_Preconditioner.hpp the way src/main.cpp:330-372 drives HYMLS::Preconditioner, on a
// serial stand-in Epetra_CrsMatrix whose row map is NOT in GID order.  Usage:
//   adapter_main <nx> <levels> <out.bin>       2D Stokes-C (nx x nx, dof 3) -- writes X (2 columns, GID order)
// exit code 0: ok, 3: the library refused because there is no CUDA device (expected on a CPU-only host)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <numeric>
#include "HYMLS_B200_Preconditioner.hpp"

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  const int levels = atoi(argv[2]);
  // matrix from a file written by the test: n, nnz, rowptr(int32), col(int32), val(f64)
  std::ifstream in(argv[1], std::ios::binary);
  int n = 0, nnz = 0, nx = 0;
  in.read((char*)&n, 4); in.read((char*)&nnz, 4); in.read((char*)&nx, 4);
  std::vector<int> ptr(n + 1), col(nnz);
  std::vector<double> val(nnz);
  in.read((char*)ptr.data(), 4 * (n + 1)); in.read((char*)col.data(), 4 * nnz); in.read((char*)val.data(), 8 * nnz);
  // a row map that is a fixed permutation of the GIDs (local row i holds GID perm[i]); columns keep GID order
  std::vector<long long> perm(n), ident(n);
  std::iota(ident.begin(), ident.end(), 0LL);
  for (int i = 0; i < n; ++i) perm[i] = ((long long)i * 7 + 3) % n;  // 7 is coprime with the test sizes
  std::vector<int> pptr(n + 1, 0), pcol;
  std::vector<double> pval;
  for (int i = 0; i < n; ++i) {
    const int g = (int)perm[i];
    for (int e = ptr[g]; e < ptr[g + 1]; ++e) { pcol.push_back(col[e]); pval.push_back(val[e]); }
    pptr[i + 1] = (int)pcol.size();
  }
  Epetra_SerialComm comm;
  Epetra_Map rowMap(n, n, perm.data(), 0, comm), colMap(n, n, ident.data(), 0, comm);
  Teuchos::RCP<const Epetra_RowMatrix> K(new Epetra_CrsMatrix(rowMap, colMap, pptr, pcol, pval));
  Teuchos::RCP<Teuchos::ParameterList> params(new Teuchos::ParameterList("HYMLS"));
  params->sublist("Problem").set("Equations", "Stokes-C").set("Dimension", 2).set("nx", nx).set("ny", nx).set("nz", 1);
  params->sublist("Preconditioner").set("Separator Length", 4).set("Number of Levels", 7);  // corrected below
  Teuchos::RCP<Epetra_Vector> tv(new Epetra_Vector(rowMap));
  {
    std::ifstream tin(argv[3], std::ios::binary);
    std::vector<double> t(n);
    tin.read((char*)t.data(), 8 * n);
    for (int i = 0; i < n; ++i) (*tv)[0][i] = t[perm[i]];
  }
  try {
    HYMLS::B200Preconditioner P(K, params, tv);
    Teuchos::ParameterList upd("HYMLS");
    upd.sublist("Preconditioner").set("Number of Levels", levels);
    if (P.SetParameters(upd) != 0) { fprintf(stderr, "SetParameters: %s\n", hymls_b200_last_error()); return 1; }
    Ifpack_Preconditioner& ifp = P;  // used through the Ifpack interface, like Belos / NOX do
    if (ifp.Initialize() != 0) { fprintf(stderr, "Initialize: %s\n", hymls_b200_last_error()); return 1; }
    if (ifp.Compute() != 0) { fprintf(stderr, "Compute: %s\n", hymls_b200_last_error()); return 1; }
    Epetra_MultiVector B(rowMap, 2), X(rowMap, 2);
    for (int i = 0; i < n; ++i) {
      B[0][i] = std::sin(0.37 * (double)perm[i]);
      B[1][i] = 1.0 / (1.0 + (double)perm[i]);
    }
    if (ifp.ApplyInverse(B, X) != 0) { fprintf(stderr, "ApplyInverse: %s\n", hymls_b200_last_error()); return 1; }
    if (ifp.Apply(B, X) != -1 || P.SetUseTranspose(true) != -1) return 1;  // as the reference
    std::vector<double> out((size_t)2 * n);
    for (int j = 0; j < 2; ++j)
      for (int i = 0; i < n; ++i) out[(size_t)j * n + perm[i]] = X[j][i];
    std::ofstream of(argv[4], std::ios::binary);
    of.write((const char*)out.data(), 8 * out.size());
    printf("adapter ok: %s, NumCompute %d, NumApplyInverse %d\n", P.Label(), ifp.NumCompute(), ifp.NumApplyInverse());
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return strstr(e.what(), "needs a CUDA device") ? 3 : 1;
  }
}
