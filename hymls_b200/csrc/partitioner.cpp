#include "partitioner.hpp"

#include <algorithm>
#include <cmath>
#include <unordered_map>

#include "../../include/hymls_b200.h"

namespace hymls {

static const int X_PERIO = 1, Y_PERIO = 2, Z_PERIO = 4;

static Error argError(const std::string& m) { return Error(HYMLS_B200_ERR_ARG, m); }

CartesianPartitioner::CartesianPartitioner(ParameterList& params, int level, int nprocs, int mypid)
    : level_(level), nprocsComm_(nprocs), mypid_(mypid) {
  setParameters(params);
}

// src/HYMLS_BasePartitioner.cpp:31-319
void CartesianPartitioner::setParameters(ParameterList& params) {
  ParameterList& prob = params.sublist("Problem");
  ParameterList& prec = params.sublist("Preconditioner");

  dim_ = prob.get("Dimension", 3);
  int pvar = -1;
  nx_ = prob.get("nx", -1);
  ny_ = prob.get("ny", nx_);
  nz_ = prob.get("nz", dim_ > 2 ? nx_ : 1);
  if (nx_ == -1) throw argError("You must presently specify nx, ny (and possibly nz) in the 'Problem' sublist");

  bool xp = prob.get("x-periodic", false);
  bool yp = dim_ > 1 ? prob.get("y-periodic", false) : false;
  bool zp = dim_ > 2 ? prob.get("z-periodic", false) : false;
  int perio = (xp ? X_PERIO : 0) | (yp ? Y_PERIO : 0) | (zp ? Z_PERIO : 0);
  perio_ = prob.get("Periodicity", perio);

  sx_ = -1;
  sy_ = -1;
  sz_ = nz_ > 1 ? -1 : 1;
  if (prec.isParameter("Separator Length (x)")) sx_ = prec.get("Separator Length (x)", sx_);
  if (prec.isParameter("Separator Length (y)")) sy_ = prec.get("Separator Length (y)", sy_);
  if (prec.isParameter("Separator Length (z)")) sz_ = prec.get("Separator Length (z)", sz_);
  if (sx_ == -1) sx_ = prec.get("Separator Length", 4);
  if (sy_ == -1) sy_ = prec.get("Separator Length", sx_);
  if (sz_ == -1) sz_ = prec.get("Separator Length", sx_);
  if (sx_ <= 1) throw argError("Separator Length not set correctly");

  cx_ = -1;
  cy_ = -1;
  cz_ = nz_ > 1 ? -1 : 1;
  if (prec.isParameter("Coarsening Factor (x)")) cx_ = prec.get("Coarsening Factor (x)", cx_);
  if (prec.isParameter("Coarsening Factor (y)")) cy_ = prec.get("Coarsening Factor (y)", cy_);
  if (prec.isParameter("Coarsening Factor (z)")) cz_ = prec.get("Coarsening Factor (z)", cz_);
  if (cx_ == -1) cx_ = prec.get("Coarsening Factor", sx_);
  if (cy_ == -1) cy_ = prec.get("Coarsening Factor", cx_);
  if (cz_ == -1) cz_ = prec.get("Coarsening Factor", cx_);
  if (cx_ <= 1) throw argError("Coarsening Factor not set correctly");

  rx_ = ry_ = rz_ = -1;
  const std::string atLevel = "Retain Nodes at Level " + std::to_string(level_);
  int* r[3] = {&rx_, &ry_, &rz_};
  const char* dn[3] = {" (x)", " (y)", " (z)"};
  for (int d = 0; d < 3; ++d) {
    if (prec.isParameter(std::string("Retain Nodes") + dn[d])) *r[d] = prec.get(std::string("Retain Nodes") + dn[d], *r[d]);
    if (prec.isParameter(atLevel + dn[d])) *r[d] = prec.get(atLevel + dn[d], *r[d]);
  }
  for (int d = 0; d < 3; ++d) {
    if (*r[d] == -1 && prec.isParameter(atLevel)) *r[d] = prec.get(atLevel, *r[d]);
    if (*r[d] == -1) *r[d] = prec.get("Retain Nodes", *r[d]);
  }

  linkRetained_ = prec.get("Eliminate Retained Nodes Together", true);
  linkVelocities_ = prec.get("Eliminate Velocities Together", true);

  if (prob.isParameter("Equations")) {
    std::string eqn = prob.get("Equations", "Undefined Problem");
    bool isComplex = prob.get("Complex Arithmetic", false);
    int factor = isComplex ? 2 : 1;
    if (eqn == "Laplace") {
      if (!isComplex) {
        prob.get("Degrees of Freedom", 1);
        prob.sublist("Variable 0").get("Variable Type", "Laplace");
      } else {
        prob.get("Degrees of Freedom", 2);
        prob.sublist("Variable 0").get("Variable Type", "Laplace");
        prob.sublist("Variable 1").get("Variable Type", "Laplace");
      }
    } else if (eqn.rfind("Stokes", 0) == 0 || eqn == "Bous-C") {
      if (eqn == "Bous-C") {
        prob.get("Degrees of Freedom", dim_ + 2);
        pvar = prob.get("Pressure Variable", dim_ + 1);
      } else {
        prob.get("Degrees of Freedom", dim_ + 1);
        pvar = prob.get("Pressure Variable", dim_);
      }
      dof_ = prob.get("Degrees of Freedom", 1);
      for (int i = 0; i < dim_ * factor; ++i)
        prob.sublist("Variable " + std::to_string(i)).get("Variable Type", "Velocity");
      for (int i = pvar * factor; i < pvar * factor + factor; ++i)
        prob.sublist("Variable " + std::to_string(i)).get("Variable Type", "Pressure");
      for (int i = 0; i < dof_; ++i)
        if (!prob.isSublist("Variable " + std::to_string(i)))
          prob.sublist("Variable " + std::to_string(i)).get("Variable Type", "Laplace");
      if (eqn == "Stokes-B" || eqn == "Stokes-L" || eqn == "Stokes-T") {
        if (isComplex) throw argError("complex Stokes-B not implemented");
        retainPressures_ = prob.get("Retained Pressure Nodes", 2);
        if (prec.get("Fix Pressure Level", true)) {
          prec.get("Fix GID 1", factor * pvar);
          prec.get("Fix GID 2", factor * dof_ + factor * pvar);
        }
      } else {
        if (prec.get("Fix Pressure Level", true)) {
          prec.get("Fix GID 1", factor * pvar);
          if (isComplex) prec.get("Fix GID 2", factor * pvar + 1);
        }
        retainPressures_ = prob.get("Retained Pressure Nodes", 1);
      }
    } else {
      throw argError("'Equations' parameter not recognized");
    }
  }
  if (!prob.isParameter("Degrees of Freedom"))
    throw argError("At this point, the 'Problem' sublist must contain 'Degrees of Freedom'");
  dof_ = prob.get("Degrees of Freedom", 1);
  retainPressures_ = prob.get("Retained Pressure Nodes", 1);

  variableType_.assign(dof_, VT_V);
  int pcount = 0, vcount = 0;
  for (int i = 0; i < dof_; ++i) {
    std::string vt = prob.sublist("Variable " + std::to_string(i)).get("Variable Type", "Laplace");
    if (vt == "Laplace") {
      variableType_[i] = VT_V;
    } else if (vt == "Velocity U" || (vt == "Velocity" && vcount == 0)) {
      variableType_[i] = VT_U;
      vcount++;
    } else if (vt == "Velocity V" || (vt == "Velocity" && vcount == 1)) {
      variableType_[i] = VT_V;
      vcount++;
    } else if (vt == "Velocity W" || (vt == "Velocity" && vcount == 2)) {
      variableType_[i] = VT_W;
      vcount++;
    } else if (vt == "Pressure") {
      pvar = i;
      variableType_[i] = VT_PRESSURE;
      pcount++;
    } else if (vt == "Interior") {
      variableType_[i] = VT_INTERIOR;
    } else {
      throw argError("Variable type " + vt + " does not exist");
    }
  }
  if (pcount > 1) throw argError("Can only have one 'Pressure' variable");
  prob.get("Pressure Variable", pvar);
  pvar_ = pvar;
  bgrid_ = prec.get("B-Grid Transform", false);
  // extension (DESIGN.md "Deviations"): default false == reference behaviour
  linkTubePressures_ = prec.get("Eliminate Tube Pressures With Velocities", false);
}

// src/HYMLS_BasePartitioner.cpp:321-346
void CartesianPartitioner::setNextLevelParameters(ParameterList& params) const {
  ParameterList& prec = params.sublist("Preconditioner");
  int nsx = sx_ * cx_, nsy = sy_ * cy_, nsz = sz_ * cz_;
  if (prec.isParameter("Separator Length (x)")) {
    prec.set("Separator Length (x)", nsx);
    prec.set("Separator Length (y)", nsy);
    prec.set("Separator Length (z)", nsz);
  } else {
    prec.set("Separator Length", nsx);
  }
  if (prec.isParameter("Coarsening Factor (x)")) {
    prec.set("Coarsening Factor (x)", cx_);
    prec.set("Coarsening Factor (y)", cy_);
    prec.set("Coarsening Factor (z)", cz_);
  } else {
    prec.set("Coarsening Factor", cx_);
  }
}

// src/HYMLS_CartesianPartitioner.cpp:80-121
void CartesianPartitioner::subdomainPosition(int sd, int sx, int sy, int sz, int& x, int& y, int& z) const {
  int npx = (nx_ - 1) / sx + 1, npy = (ny_ - 1) / sy + 1, npz = (nz_ - 1) / sz + 1;
  x = (sd % npx) * sx;
  y = ((sd / npx) % npy) * sy;
  z = ((sd / npx / npy) % npz) * sz;
}
int CartesianPartitioner::subdomainId(int sx, int sy, int sz, int x, int y, int z) const {
  int npx = (nx_ - 1) / sx + 1, npy = (ny_ - 1) / sy + 1;
  return (z / sz * npy + y / sy) * npx + x / sx;
}
int CartesianPartitioner::numGlobalParts(int sx, int sy, int sz) const {
  return ((nx_ - 1) / sx + 1) * ((ny_ - 1) / sy + 1) * ((nz_ - 1) / sz + 1);
}
int CartesianPartitioner::pid(gidx gid) const {
  gidx rem = gid / dof_;
  int i = (int)(rem % nx_);
  rem /= nx_;
  int j = (int)(rem % ny_);
  rem /= ny_;
  int k = (int)(rem % nz_);
  return pidMap_[subdomainId(sx_, sy_, sz_, i, j, k)];
}

static int findCoarseningFactor(int cx) {  // src/HYMLS_BasePartitioner.cpp:348-359
  int b = 1;
  while (b < cx) {
    for (int p = 0; p < cx; ++p)
      if (std::pow((double)b, (double)p) == (double)cx) return b;
    b += 1;
  }
  return cx;
}

// src/HYMLS_BasePartitioner.cpp:361-586
void CartesianPartitioner::createPidMap() {
  int sx = sx_, sy = sy_, sz = sz_;
  int nparts = numGlobalParts(sx, sy, sz);
  const int P = nprocsComm_;
  if (P == 1 || nparts == 1) {
    nprocs_ = 1;
    pidMap_.assign(nparts, 0);
    return;
  }
  pidMap_.assign(nparts, -1);
  std::vector<std::vector<int>> pidGroups(nparts);
  std::vector<int> sdPidNum(nparts, 0);
  int cx = findCoarseningFactor(cx_), cy = findCoarseningFactor(cy_), cz = findCoarseningFactor(cz_);
  while (sx < nx_ || sy < ny_ || sz < nz_) {
    sx *= cx;
    sy *= cy;
    if (nz_ > 1) sz *= cz;
  }
  int sx2 = sx, sy2 = sy, sz2 = sz;
  auto wrap = [&](int& x, int& y, int& z) {
    x = (x % nx_ + nx_) % nx_;
    y = (y % ny_ + ny_) % ny_;
    z = (z % nz_ + nz_) % nz_;
  };
  nprocs_ = 0;
  for (int j = 0; j < 1000; ++j) {
    nparts = numGlobalParts(sx, sy, sz);
    int prevNprocs = nprocs_;
    std::vector<std::vector<int>> prevGroups = pidGroups;
    for (int i = 0; i < nparts; ++i) {
      int x, y, z;
      subdomainPosition(i, sx, sy, sz, x, y, z);
      wrap(x, y, z);
      int sd = subdomainId(sx_, sy_, sz_, x, y, z);
      if (pidGroups[sd].empty()) pidGroups[sd].push_back(nprocs_++);
    }
    if (nprocs_ > P) {
      nprocs_ = prevNprocs;
      pidGroups = prevGroups;
      break;
    }
    sx2 = sx;
    sy2 = sy;
    sz2 = sz;
    sx /= cx;
    sy /= cy;
    if (nz_ > 1) sz /= cz;
    if (sx < sx_ || sy < sy_ || sz < sz_) {
      sx = sx2;
      sy = sy2;
      sz = sz2;
      break;
    }
  }
  nparts = numGlobalParts(sx_, sy_, sz_);
  for (int j = 0; j < 1000; ++j) {
    if (nprocs_ >= P) break;
    for (int sd = 0; sd < nparts; ++sd) {
      if (nprocs_ >= P) break;
      if (!pidGroups[sd].empty()) pidGroups[sd].push_back(nprocs_++);
    }
  }
  nparts = numGlobalParts(sx, sy, sz);
  for (int i = 0; i < nparts; ++i) {
    int x, y, z;
    subdomainPosition(i, sx, sy, sz, x, y, z);
    wrap(x, y, z);
    int sd = subdomainId(sx_, sy_, sz_, x, y, z);
    if (pidMap_[sd] != -1) continue;
    int sd2 = subdomainId(sx2, sy2, sz2, x, y, z);
    subdomainPosition(sd2, sx2, sy2, sz2, x, y, z);
    wrap(x, y, z);
    sd2 = subdomainId(sx_, sy_, sz_, x, y, z);
    if (pidGroups[sd2].empty()) throw argError("CreatePIDMap: invalid subdomain index");
    pidMap_[sd] = pidGroups[sd2][sdPidNum[sd2]++ % pidGroups[sd2].size()];
  }
  nparts = numGlobalParts(sx_, sy_, sz_);
  for (int i = 0; i < nparts; ++i) {
    if (pidMap_[i] != -1) continue;
    int x, y, z;
    subdomainPosition(i, sx_, sy_, sz_, x, y, z);
    wrap(x, y, z);
    int sd = subdomainId(sx_, sy_, sz_, x, y, z);
    if (pidMap_[sd] != -1) {
      pidMap_[i] = pidMap_[sd];
      continue;
    }
    sd = subdomainId(sx, sy, sz, x, y, z);
    subdomainPosition(sd, sx, sy, sz, x, y, z);
    wrap(x, y, z);
    sd = subdomainId(sx_, sy_, sz_, x, y, z);
    if (pidMap_[sd] == -1) throw argError("CreatePIDMap: invalid subdomain index");
    pidMap_[i] = pidMap_[sd];
  }
  std::vector<int> tmp(pidMap_);
  std::sort(tmp.begin(), tmp.end());
  nprocs_ = (int)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
}

void CartesianPartitioner::partition() {
  createPidMap();
  sdMap_.clear();
  int nparts = numGlobalParts(sx_, sy_, sz_);
  for (int sd = 0; sd < nparts; ++sd)
    if (pidMap_[sd] == mypid_) sdMap_.push_back(sd);
}

// src/HYMLS_CartesianPartitioner.cpp:224-263
static int startAndEnd(int pos, int idx, int idxMax, int dim, int mx, bool perio, int& type, int& start,
                       int& end) {
  int len = std::max((mx + idxMax - 1) / idxMax, 1);
  if (idx == idxMax)
    type = 2;
  else if (idx >= 0)
    type = 1;
  else
    type = 0;
  start = idx;
  if (idx == idxMax)
    start = mx;
  else if (idx > 0)
    start = std::min(len * idx, mx);
  end = start + 1;
  if (type == 1) end = std::min(len * (idx + 1), mx);
  if (!perio) {
    if (pos == 0 && idx == -1) return 1;
    if (pos + mx + 1 == dim) {
      if (idx == idxMax) return 1;
      if (idx == idxMax - 1) end += 1;
    }
  }
  if (start == end) return 1;
  return 0;
}

// src/HYMLS_CartesianPartitioner.cpp:265-408
void CartesianPartitioner::getGroups(int localSd, std::vector<gidx>& interior,
                                     std::vector<SepGroup>& groups) const {
  interior.clear();
  groups.clear();
  std::vector<gidx> retained;
  int gsd = sdMap_[localSd];
  int xpos, ypos, zpos;
  subdomainPosition(gsd, sx_, sy_, sz_, xpos, ypos, zpos);
  int xmax = std::min(nx_ - xpos - 1, sx_ - 1);
  int ymax = std::min(ny_ - ypos - 1, sy_ - 1);
  int zmax = std::min(nz_ - zpos - 1, sz_ - 1);
  if (xmax == 0 || ymax == 0 || (zmax == 0 && nz_ > 1)) throw argError("Can't have a subdomain of size 1");
  int iMax = rx_ > 1 ? rx_ : 1, jMax = ry_ > 1 ? ry_ : 1, kMax = rz_ > 1 ? rz_ : 1;

  for (int kidx = -1; kidx <= kMax; ++kidx) {
    bool kint = kidx >= 0 && kidx < kMax;
    int ktype, kstart, kend;
    if (startAndEnd(zpos, kidx, kMax, nz_, zmax, perio_ & Z_PERIO, ktype, kstart, kend)) continue;
    for (int jidx = -1; jidx <= jMax; ++jidx) {
      bool jint = jidx >= 0 && jidx < jMax;
      int jtype, jstart, jend;
      if (startAndEnd(ypos, jidx, jMax, ny_, ymax, perio_ & Y_PERIO, jtype, jstart, jend)) continue;
      for (int iidx = -1; iidx <= iMax; ++iidx) {
        bool iint = iidx >= 0 && iidx < iMax;
        int itype, istart, iend;
        if (startAndEnd(xpos, iidx, iMax, nx_, xmax, perio_ & X_PERIO, itype, istart, iend)) continue;
        for (int d = 0; d < dof_; ++d) {
          const int vt = variableType_[d];
          // destination: -1 interior, otherwise index of the group in `groups`
          int dst = -1, dst2 = -2;
          if ((vt == VT_PRESSURE || vt == VT_INTERIOR) && (iidx == -1 || jidx == -1 || kidx == -1)) {
            continue;
          } else if ((iint && jint && kint) || vt == VT_INTERIOR ||
                     (vt == VT_PRESSURE &&
                      ((iint && jint) || (iint && kint) || (jint && kint) || retainPressures_ > 1))) {
            dst = -1;
          } else {
            int type = -1000;
            if (linkRetained_) type = 2 * dof_ * (itype + 3 * (jtype + 3 * ktype));
            bool isVel = vt == VT_U || vt == VT_V || vt == VT_W;
            if (!((linkVelocities_ && isVel) || (linkTubePressures_ && vt == VT_PRESSURE))) type += 2 * d;
            groups.emplace_back();
            groups.back().type = type;
            dst = (int)groups.size() - 1;
            if (bgrid_) {
              groups.emplace_back();
              groups.back().type = type + 1;
              dst2 = (int)groups.size() - 1;
            }
          }
          for (int k = kstart; k < kend; ++k)
            for (int j = jstart; j < jend; ++j)
              for (int i = istart; i < iend; ++i) {
                gidx gid = d + (gidx)((i + xpos + nx_) % nx_) * dof_ +
                           (gidx)((j + ypos + ny_) % ny_) * nx_ * dof_ +
                           (gidx)((k + zpos + nz_) % nz_) * nx_ * ny_ * dof_;
                if (vt == VT_PRESSURE && i >= 0 && j >= 0 && k >= 0 && (int)retained.size() < retainPressures_) {
                  retained.push_back(gid);
                } else if (dst2 >= 0 && (i + xpos + j + ypos) % 2) {
                  groups[dst2].nodes.push_back(gid);
                } else if (dst < 0) {
                  interior.push_back(gid);
                } else {
                  groups[dst].nodes.push_back(gid);
                }
              }
        }
      }
    }
  }
  groups.erase(std::remove_if(groups.begin(), groups.end(), [](const SepGroup& g) { return g.nodes.empty(); }),
               groups.end());
  for (gidx g : retained) {
    groups.emplace_back();
    groups.back().type = -1;
    groups.back().nodes.push_back(g);
  }
}

std::vector<std::vector<int>> linkGroups(const std::vector<int>& types) {
  std::vector<std::vector<int>> out;
  for (int gi = 0; gi < (int)types.size(); ++gi) {
    bool found = false;
    if (types[gi] >= 0) {
      for (auto& lg : out)
        if (types[lg[0]] == types[gi]) {
          lg.push_back(gi);
          found = true;
          break;
        }
    }
    if (!found) out.push_back(std::vector<int>(1, gi));
  }
  return out;
}

// OverlappingPartitioner::DetectSeparators + HierarchicalMap::FillComplete (single rank view)
void buildHierarchicalMap(const CartesianPartitioner& part, const std::vector<char>& present,
                          HierarchicalMap& H) {
  H = HierarchicalMap();
  const int nsd = part.numLocalParts();
  H.nsd = nsd;
  H.intPtr.assign(1, 0);
  H.sdGrpPtr.assign(1, 0);
  H.grpPtr.assign(1, 0);
  H.uniqPtr.assign(1, 0);
  const bool filter = !present.empty();
  std::unordered_map<gidx, int> uniqueByFirst;
  uniqueByFirst.reserve((size_t)nsd * 32);
  std::vector<gidx> interior;
  std::vector<SepGroup> groups;
  for (int sd = 0; sd < nsd; ++sd) {
    part.getGroups(sd, interior, groups);
    std::sort(interior.begin(), interior.end());
    for (gidx g : interior)
      if (!filter || present[g]) {
        H.intGid.push_back(g);
        H.overlappingGid.push_back(g);
      }
    H.intPtr.push_back((int64_t)H.intGid.size());
    for (auto& grp : groups) {
      std::sort(grp.nodes.begin(), grp.nodes.end());
      size_t before = H.grpGid.size();
      for (gidx g : grp.nodes)
        if (!filter || present[g]) H.grpGid.push_back(g);
      if (H.grpGid.size() == before) continue;  // empty after filtering: removed (:241-244)
      H.grpPtr.push_back((int64_t)H.grpGid.size());
      H.grpType.push_back(grp.type);
      gidx first = H.grpGid[before];
      auto it = uniqueByFirst.find(first);
      int u;
      if (it == uniqueByFirst.end()) {
        u = (int)H.uniqOwnerSd.size();
        uniqueByFirst.emplace(first, u);
        H.uniqOwnerSd.push_back(sd);
        H.uniqType.push_back(grp.type);
        for (size_t q = before; q < H.grpGid.size(); ++q) {
          H.sepGid.push_back(H.grpGid[q]);
          H.overlappingGid.push_back(H.grpGid[q]);
        }
        H.uniqPtr.push_back((int64_t)H.sepGid.size());
      } else {
        u = it->second;
        // the reference identifies groups by their first GID only (:261-271); a mismatch in the
        // node list would make its maps inconsistent, so we refuse it.
        int64_t len = H.uniqPtr[u + 1] - H.uniqPtr[u];
        if (len != (int64_t)(H.grpGid.size() - before))
          throw Error(HYMLS_B200_ERR_ARG, "separator group seen with two different node lists");
      }
      H.grpUnique.push_back(u);
    }
    H.sdGrpPtr.push_back((int64_t)H.grpType.size());
  }
}

}  // namespace hymls
