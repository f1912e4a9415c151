#pragma once
#include "Epetra_Map.h"
class Epetra_RowMatrix {
 public:
  virtual ~Epetra_RowMatrix() {}
  virtual int NumMyRows() const = 0;
  virtual const Epetra_Comm& Comm() const = 0;
};
