#pragma once
#include <vector>
#include "Epetra_Map.h"
class Epetra_MultiVector {
 public:
  Epetra_MultiVector(const Epetra_BlockMap& map, int numVectors)
      : n_(map.NumMyElements()), nv_(numVectors), v_((size_t)n_ * numVectors, 0.0) {}
  double* Values() const { return const_cast<double*>(v_.data()); }
  int Stride() const { return n_; }
  int NumVectors() const { return nv_; }
  int MyLength() const { return n_; }
  double* operator[](int j) { return v_.data() + (size_t)j * n_; }
  const double* operator[](int j) const { return v_.data() + (size_t)j * n_; }
 private:
  int n_, nv_;
  std::vector<double> v_;
};
class Epetra_Vector : public Epetra_MultiVector {
 public:
  explicit Epetra_Vector(const Epetra_BlockMap& map) : Epetra_MultiVector(map, 1) {}
};
