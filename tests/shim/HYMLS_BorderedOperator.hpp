#pragma once
// interface of src/HYMLS_BorderedOperator.hpp:18-63 (the reference tree's header)
#include "Teuchos_RCP.hpp"
class Epetra_MultiVector;
class Epetra_SerialDenseMatrix;
namespace HYMLS {
class BorderedOperator {
 public:
  virtual ~BorderedOperator() {}
  virtual int SetBorder(Teuchos::RCP<const Epetra_MultiVector> V, Teuchos::RCP<const Epetra_MultiVector> W,
                        Teuchos::RCP<const Epetra_SerialDenseMatrix> C) = 0;
  virtual int Apply(const Epetra_MultiVector& X, const Epetra_SerialDenseMatrix& S, Epetra_MultiVector& Y,
                    Epetra_SerialDenseMatrix& T) const = 0;
  virtual int ApplyInverse(const Epetra_MultiVector& X, const Epetra_SerialDenseMatrix& S, Epetra_MultiVector& Y,
                           Epetra_SerialDenseMatrix& T) const = 0;
};
}  // namespace HYMLS
