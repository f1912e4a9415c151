#pragma once
#include <vector>
#include "Epetra_RowMatrix.h"
// serial stand-in: CSR with LOCAL column indices and a column map, like a FillComplete'd Epetra_CrsMatrix
class Epetra_CrsMatrix : public Epetra_RowMatrix {
 public:
  Epetra_CrsMatrix(const Epetra_Map& rowMap, const Epetra_Map& colMap, const std::vector<int>& ptr,
                   const std::vector<int>& lcol, const std::vector<double>& val)
      : rowMap_(rowMap), colMap_(colMap), ptr_(ptr), col_(lcol), val_(val) {}
  int NumMyRows() const { return rowMap_.NumMyElements(); }
  int NumMyNonzeros() const { return (int)col_.size(); }
  long long NumGlobalRows64() const { return rowMap_.NumGlobalElements64(); }
  int ExtractMyRowView(int row, int& numEntries, double*& values, int*& indices) const {
    numEntries = ptr_[row + 1] - ptr_[row];
    values = const_cast<double*>(val_.data()) + ptr_[row];
    indices = const_cast<int*>(col_.data()) + ptr_[row];
    return 0;
  }
  long long GCID64(int lcid) const { return colMap_.GID64(lcid); }
  const Epetra_Map& RowMap() const { return rowMap_; }
  const Epetra_Map& ColMap() const { return colMap_; }
  const Epetra_Map& OperatorDomainMap() const { return rowMap_; }
  const Epetra_Map& OperatorRangeMap() const { return rowMap_; }
  const Epetra_Comm& Comm() const { return rowMap_.Comm(); }
  std::vector<double>& Values() { return val_; }
 private:
  Epetra_Map rowMap_, colMap_;
  std::vector<int> ptr_, col_;
  std::vector<double> val_;
};
