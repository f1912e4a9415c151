// Host-side index-set construction (bit-exact with the reference):
//   BasePartitioner::SetParameters / CreatePIDMap      src/HYMLS_BasePartitioner.cpp:31-319, 361-586
//   CartesianPartitioner::GetGroups                    src/HYMLS_CartesianPartitioner.cpp:224-408
//   HierarchicalMap::FillComplete / LinkSeparators     src/HYMLS_HierarchicalMap.cpp:120-285
//   OverlappingPartitioner (levels, SpawnNextLevel)    src/HYMLS_OverlappingPartitioner.cpp:31-159
// Output is a flat, device-friendly description (CSR-like arrays), not a tree of objects.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "params.hpp"

namespace hymls {

typedef int64_t gidx;

enum VarType { VT_U = 0, VT_V = 1, VT_W = 2, VT_PRESSURE = 3, VT_INTERIOR = 4 };

struct SepGroup {
  int type = -1;
  std::vector<gidx> nodes;
};

class CartesianPartitioner {
 public:
  CartesianPartitioner(ParameterList& params, int level, int nprocs = 1, int mypid = 0);
  virtual ~CartesianPartitioner() {}

  virtual void partition();  // CreatePIDMap + CreateSubdomainMap
  int numLocalParts() const { return (int)sdMap_.size(); }
  int numGlobalParts() const { return numGlobalParts(sx_, sy_, sz_); }
  int globalSubdomain(int localSd) const { return sdMap_[localSd]; }
  const std::vector<int>& pidMap() const { return pidMap_; }
  int numActiveProcs() const { return nprocs_; }

  // interior nodes and separator groups in the reference's emission order (not yet sorted)
  virtual void getGroups(int localSd, std::vector<gidx>& interior, std::vector<SepGroup>& groups) const;
  void setNextLevelParameters(ParameterList& params) const;

  virtual int subdomainId(int sx, int sy, int sz, int x, int y, int z) const;
  virtual int subdomainPosition(int sd, int sx, int sy, int sz, int& x, int& y, int& z) const;
  int pid(gidx gid) const;

  int nx() const { return nx_; }
  int ny() const { return ny_; }
  int nz() const { return nz_; }
  int dof() const { return dof_; }
  int sx() const { return sx_; }
  int pvar() const { return pvar_; }
  gidx numGlobalNodes() const { return (gidx)nx_ * ny_ * nz_ * dof_; }

 protected:
  void setParameters(ParameterList& params);
  virtual int numGlobalParts(int sx, int sy, int sz) const;
  void createPidMap();
  // fine-grid subdomain that contains the (periodically wrapped) origin of brick `b` of the brick grid (bx,by,bz)
  int brickAnchor(int b, int bx, int by, int bz) const;

  int level_, nprocsComm_, mypid_;
  int dim_ = 3, nx_ = 0, ny_ = 0, nz_ = 0, dof_ = 1, perio_ = 0, pvar_ = -1;
  int sx_ = 0, sy_ = 0, sz_ = 0, cx_ = 0, cy_ = 0, cz_ = 0, rx_ = -1, ry_ = -1, rz_ = -1;
  int retainPressures_ = 1;
  bool linkRetained_ = true, linkVelocities_ = true, bgrid_ = false, linkTubePressures_ = false;
  std::vector<int> variableType_;
  std::vector<int> pidMap_;
  std::vector<int> sdMap_;
  int nprocs_ = 1;
};

// Skew Cartesian partitioner (behaviour of src/HYMLS_SkewCartesianPartitioner.cpp, formulated geometrically):
// the subdomains are the unit cubes of side sx of the sheared lattice coordinates
//     s = x + y,   d = x - y,   t = x - y + z,
// and the node set ("template") a subdomain touches is, per variable type, the set of lattice points of a
// polytope bounded by planes of constant s, d, t and z (partitioner.cpp: SkewShape).  Same BasePartitioner
// machinery (parameters, CreatePIDMap) as the Cartesian one.
class SkewCartesianPartitioner : public CartesianPartitioner {
 public:
  SkewCartesianPartitioner(ParameterList& params, int level, int nprocs = 1, int mypid = 0)
      : CartesianPartitioner(params, level, nprocs, mypid) {}
  void partition() override;
  void getGroups(int localSd, std::vector<gidx>& interior, std::vector<SepGroup>& groups) const override;
  int subdomainId(int sx, int sy, int sz, int x, int y, int z) const override;
  int subdomainPosition(int sd, int sx, int sy, int sz, int& x, int& y, int& z) const override;

  // one node of the subdomain template: offset from the subdomain position, variable, class
  // (0 = touched by this subdomain only, k >= 1 = k-th distinct set of neighbouring subdomains met in scan order)
  struct TemplateNode { int dx, dy, dz, var, cls; };

 protected:
  int numGlobalParts(int sx, int sy, int sz) const override;

 private:
  void classifyTemplate();
  std::vector<TemplateNode> tmpl_;      // sorted by (cls, var), scan order (z, y, x) inside
  std::vector<int64_t> clsVarPtr_;      // (ncls * dof + 1): ranges of tmpl_ per (class, variable)
  std::vector<TemplateNode> innerScan_; // class 0 in scan order (z, y, x, variable)
  int ncls_ = 0;
};

// factory: "Cartesian" | "Skew Cartesian" (OverlappingPartitioner::Partition, :96-119)
CartesianPartitioner* makePartitioner(ParameterList& params, int level, int nprocs = 1, int mypid = 0);

// One level of the hierarchy after FillComplete, flattened.
struct HierarchicalMap {
  int nsd = 0;
  // interior nodes, per subdomain, ascending GIDs
  std::vector<int64_t> intPtr;  // nsd+1
  std::vector<gidx> intGid;
  // ALL separator groups around each subdomain, reference order
  std::vector<int64_t> sdGrpPtr;  // nsd+1 -> index into grp arrays
  std::vector<int64_t> grpPtr;    // ngrp+1 -> index into grpGid
  std::vector<gidx> grpGid;
  std::vector<int> grpType;
  std::vector<int> grpUnique;  // unique-group id of every (sd, group)
  // unique groups (first local subdomain that lists them owns them); separator ordering =
  // concatenation of unique groups in (sd, group) order  (SpawnSeparators, :470-508)
  std::vector<int64_t> uniqPtr;  // nuniq+1 -> position in separator ordering
  std::vector<int> uniqOwnerSd;
  std::vector<int> uniqType;  // type as seen by the owner
  std::vector<gidx> sepGid;   // separator map
  // overlapping (row) map: per sd interior then its unique groups (:248-275)
  std::vector<gidx> overlappingGid;

  int64_t numInterior() const { return (int64_t)intGid.size(); }
  int64_t numSeparator() const { return (int64_t)sepGid.size(); }
  int numUnique() const { return (int)uniqOwnerSd.size(); }
};

// Builds the HierarchicalMap of one level.  `present` (size = number of GIDs of the fine grid, or
// empty at level 0) flags the GIDs that exist in this level's map.
void buildHierarchicalMap(const CartesianPartitioner& part, const std::vector<char>& present,
                          HierarchicalMap& out);

// groups of one subdomain linked by equal type >= 0, first-seen order (LinkSeparators, :120-142).
// `types` are the types of the candidate groups; returns lists of indices into that array.
std::vector<std::vector<int>> linkGroups(const std::vector<int>& types);

}  // namespace hymls
