#pragma once
class Epetra_Comm {
 public:
  virtual ~Epetra_Comm() {}
  virtual int NumProc() const = 0;
  virtual int MyPID() const = 0;
};
class Epetra_SerialComm : public Epetra_Comm {
 public:
  int NumProc() const { return 1; }
  int MyPID() const { return 0; }
};
