#pragma once
#include <vector>
class Epetra_SerialDenseMatrix {
 public:
  Epetra_SerialDenseMatrix() {}
  Epetra_SerialDenseMatrix(int m, int n) { Shape(m, n); }
  int Shape(int m, int n) { m_ = m; n_ = n; a_.assign((size_t)m * n, 0.0); return 0; }
  int M() const { return m_; }
  int N() const { return n_; }
  double& operator()(int i, int j) { return a_[i + (size_t)j * m_]; }
  const double& operator()(int i, int j) const { return a_[i + (size_t)j * m_]; }
 private:
  int m_ = 0, n_ = 0;
  std::vector<double> a_;
};
