#pragma once
#include <unordered_map>
#include <vector>
#include "Epetra_Comm.h"
class Epetra_BlockMap {
 public:
  Epetra_BlockMap(long long numGlobal, int numMy, const long long* gids, int /*indexBase*/, const Epetra_Comm& comm)
      : n_(numGlobal), gids_(gids, gids + numMy), comm_(&comm) {
    for (int i = 0; i < numMy; ++i) lid_[gids_[i]] = i;
  }
  int NumMyElements() const { return (int)gids_.size(); }
  long long NumGlobalElements64() const { return n_; }
  long long GID64(int lid) const { return gids_[lid]; }
  int LID(long long gid) const { auto it = lid_.find(gid); return it == lid_.end() ? -1 : it->second; }
  const Epetra_Comm& Comm() const { return *comm_; }
 private:
  long long n_;
  std::vector<long long> gids_;
  std::unordered_map<long long, int> lid_;
  const Epetra_Comm* comm_;
};
class Epetra_Map : public Epetra_BlockMap {
 public:
  using Epetra_BlockMap::Epetra_BlockMap;
};
