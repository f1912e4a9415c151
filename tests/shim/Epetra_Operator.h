#pragma once
class Epetra_MultiVector;
class Epetra_Map;
class Epetra_Comm;
class Epetra_Operator {
 public:
  virtual ~Epetra_Operator() {}
  virtual int SetUseTranspose(bool UseTranspose) = 0;
  virtual int Apply(const Epetra_MultiVector& X, Epetra_MultiVector& Y) const = 0;
  virtual int ApplyInverse(const Epetra_MultiVector& X, Epetra_MultiVector& Y) const = 0;
  virtual double NormInf() const = 0;
  virtual const char* Label() const = 0;
  virtual bool UseTranspose() const = 0;
  virtual bool HasNormInf() const = 0;
  virtual const Epetra_Comm& Comm() const = 0;
  virtual const Epetra_Map& OperatorDomainMap() const = 0;
  virtual const Epetra_Map& OperatorRangeMap() const = 0;
};
