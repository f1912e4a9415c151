#include "partitioner.hpp"

#include "hostpar.hpp"

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
int CartesianPartitioner::subdomainPosition(int sd, int sx, int sy, int sz, int& x, int& y, int& z) const {
  int npx = (nx_ - 1) / sx + 1, npy = (ny_ - 1) / sy + 1, npz = (nz_ - 1) / sz + 1;
  x = (sd % npx) * sx;
  y = ((sd / npx) % npy) * sy;
  z = ((sd / npx / npy) % npz) * sz;
  return 0;
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
  // the per-subdomain group lists are independent (GetGroups is const): build and sort them in parallel, then
  // merge sequentially in subdomain order (the order decides which subdomain owns a shared group)
  std::vector<std::vector<gidx>> allInterior(nsd);
  std::vector<std::vector<SepGroup>> allGroups(nsd);
  std::vector<std::string> errors(64);
  parallelFor(nsd, [&](int64_t s0, int64_t s1, int t) {
    try {
      for (int64_t sd = s0; sd < s1; ++sd) {
        part.getGroups((int)sd, allInterior[sd], allGroups[sd]);
        std::sort(allInterior[sd].begin(), allInterior[sd].end());
        for (auto& grp : allGroups[sd]) std::sort(grp.nodes.begin(), grp.nodes.end());
      }
    } catch (const std::exception& e) {
      errors[t & 63] = e.what();
    }
  }, 16);
  for (const std::string& e : errors)
    if (!e.empty()) throw Error(HYMLS_B200_ERR_ARG, e);
  for (int sd = 0; sd < nsd; ++sd) {
    std::vector<gidx>& interior = allInterior[sd];
    std::vector<SepGroup>& groups = allGroups[sd];
    for (gidx g : interior)
      if (!filter || present[g]) {
        H.intGid.push_back(g);
        H.overlappingGid.push_back(g);
      }
    H.intPtr.push_back((int64_t)H.intGid.size());
    for (auto& grp : groups) {
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
    std::vector<gidx>().swap(interior);   // release as we go
    std::vector<SepGroup>().swap(groups);
  }
}

}  // namespace hymls

// =============================================================================================
// Skew Cartesian partitioner (src/HYMLS_SkewCartesianPartitioner.cpp)
// =============================================================================================
namespace hymls {

CartesianPartitioner* makePartitioner(ParameterList& params, int level, int nprocs, int mypid) {
  std::string method = params.sublist("Preconditioner").get("Partitioner", "Cartesian");
  if (method == "Cartesian") return new CartesianPartitioner(params, level, nprocs, mypid);
  if (method == "Skew Cartesian") return new SkewCartesianPartitioner(params, level, nprocs, mypid);
  throw Error(HYMLS_B200_ERR_ARG, "Up to now we only support Cartesian partitioning");
}

// :128-160
int SkewCartesianPartitioner::subdomainPosition(int sd, int sx, int sy, int sz, int& x, int& y, int& z) const {
  (void)sz;
  const int npx = nx_ / sx, npy = ny_ / sy;
  const int perLayer = 2 * npx * npy + npx + npy;
  const int perRow = 2 * npx + 1;
  const int Z = perLayer > 0 ? sd / perLayer : 0;
  int Y = ((sd - Z * perLayer) / perRow) * 2 - 1;
  int X = ((sd - Z * perLayer) % perRow) * 2;
  if (X >= npx * 2) {
    X -= npx * 2 + 1;
    Y += 1;
  }
  x = (X * sx) / 2;
  y = (Y * sx) / 2 + sx / 2;
  z = Z * sx;
  if (x == nx_ - sx / 2 && (perio_ & X_PERIO)) return 1;
  if (y == ny_ && (perio_ & Y_PERIO)) return 1;
  if (z == nz_ && (perio_ & Z_PERIO)) return 1;
  return 0;
}

// :162-207
int SkewCartesianPartitioner::subdomainId(int sx, int sy, int sz, int x, int y, int z) const {
  const int npx = nx_ / sx, npy = ny_ / sy, npz = nz_ / sz;
  const int dir1 = npx + 1, dir2 = npx, dir3 = 2 * npx * npy + npx + npy;
  const int xc = x / sx, yc = y / sx, zc = z / sx;
  int sd = zc * dir3 + yc * (dir2 + dir1) + xc;
  x -= xc * sx - 1;
  y -= yc * sx;
  z -= zc * sx;
  const bool front = y < sx - x;
  const bool right = y < x;
  bool below = z <= y - x;
  if (right) below = z <= sx + y - x;
  if (!front) sd += dir1;
  if (!right) sd += dir2;
  if (!below) sd += dir3;
  if (!front && right && (perio_ & X_PERIO) && xc == npx - 1) sd -= dir2;
  if (!front && !right && (perio_ & Y_PERIO) && yc == npy - 1) sd -= dir3 - dir2;
  if (!below && (perio_ & Z_PERIO) && zc == npz - 1) sd -= npz * dir3;
  return sd;
}

// :219-238
int SkewCartesianPartitioner::numGlobalParts(int sx, int sy, int sz) const {
  const int npx = nx_ / sx, npy = ny_ / sy, npz = nz_ / sz;
  const int perLayer = 2 * npx * npy + npx + npy;
  int n = perLayer;
  if (nz_ > 1) n += perLayer * npz;
  return std::max(n, 1);
}

// :240-368
void SkewCartesianPartitioner::partition() {
  if (sx_ != sy_ || (nz_ > 1 && sx_ != sz_)) throw argError("sx, sy and sz should be the same");
  if ((sx_ / 2) * 2 != sx_) throw argError("sx should be even");
  if (nx_ % sx_ || ny_ % sy_ || nz_ % sz_)
    throw argError("You are trying to partition a domain whose size is not a multiple of the subdomain size");
  createPidMap();
  sdMap_.clear();
  const int nparts = numGlobalParts(sx_, sy_, sz_);
  for (int sd = 0; sd < nparts; ++sd) {
    int i, j, k;
    if (subdomainPosition(sd, sx_, sy_, sz_, i, j, k) == 1) continue;
    i = (i % nx_ + nx_) % nx_;
    j = (j % ny_ + ny_) % ny_;
    k = (k % nz_ + nz_) % nz_;
    if (pidMap_[subdomainId(sx_, sy_, sz_, i, j, k)] == mypid_) sdMap_.push_back(sd);
  }
  buildTemplate();
  solveGroups();
}

namespace {
struct Plane45 {
  std::vector<long long> ptr, plane;
};
// :27-78: a diamond of nodes in the xy-plane, row by row
Plane45 buildPlane45(long long first, int length, long long dirX, long long dirY, int type) {
  long long left = first, right = first;
  int height = 2 * length;
  bool extra = false;
  const long long dir1 = dirY + dirX, dir2 = dirY - dirX;
  if (type == 0) {
    left -= dirX;
    height++;
    extra = true;
  } else if (type == 3) {
    height++;
    extra = true;
  }
  Plane45 P;
  P.ptr.push_back(0);
  for (int i = 0; i < height - 1; ++i) {
    for (long long j = left; j <= right; j += dirX) P.plane.push_back(j);
    P.ptr.push_back((long long)P.plane.size());
    if (i < length - 1) {
      left += dir2;
      right += dir1;
    } else if (extra && i == length - 1) {
      left += dirY;
      right += dirY;
    } else {
      left += dir1;
      right += dir2;
    }
  }
  return P;
}
}  // namespace

// getTemplate :372-565
void SkewCartesianPartitioner::buildTemplate() {
  const int sx = sx_, dof = dof_;
  const long long nx = sx * 4;
  const long long dirX = dof, dirY = dof * nx, dirZ = dof * nx * nx;
  const long long first[4] = {dof * sx / 2 + dirY + dirZ * sx, dof * sx / 2 + dirZ * sx,
                              dof * sx / 2 + dirY + dirZ * sx, dof * sx / 2 + dirY + dirZ * sx};
  const int baseLen[4] = {sx / 2, sx / 2 + 1, sx / 2 + 1, sx / 2};
  const int typeArr[4] = {VT_U, VT_V, VT_W, VT_PRESSURE};
  std::vector<std::vector<std::vector<long long>>> nodes(4);
  for (int type = 0; type < 4; ++type) {
    auto& layers = nodes[type];
    layers.assign(2 * sx + 1, std::vector<long long>());
    Plane45 P = buildPlane45(first[type], baseLen[type], dirX, dirY, type);
    layers[sx] = P.plane;
    if (nz_ <= 1) continue;
    std::vector<long long> bottom, top = P.plane, rowLen;
    for (size_t i = 0; i + 1 < P.ptr.size(); ++i) rowLen.push_back(P.ptr[i + 1] - P.ptr[i] - 1);
    std::vector<long long> active, offset;
    for (int i = 0; i < baseLen[type]; ++i) active.push_back(i);
    for (long long a : active) offset.push_back(rowLen[a]);
    for (int i = 0; i < sx; ++i) {
      for (size_t j = 0; j < active.size(); ++j) {
        const long long val = P.plane[P.ptr[active[j]] + offset[j]];
        bottom.push_back(val);
        top.erase(std::remove(top.begin(), top.end(), val), top.end());
      }
      if (typeArr[type] == VT_W) {
        if (i % 2 == 1) {
          for (long long j : top) layers[sx + i].push_back(j + i * dirZ - dirY);
          for (long long j : top) layers[sx + 1 + i].push_back(j + (i + 1) * dirZ);
        } else {
          for (long long j : bottom) layers[i].push_back(j - (sx - i) * dirZ);
          if (i > 0) {
            for (long long j : bottom) layers[i - 1].push_back(j - (sx - i + 1) * dirZ - dirY);
          } else {
            for (long long j : P.plane) layers[sx - 1].push_back(j - dirZ - dirY);
          }
        }
      } else {
        const int isP = typeArr[type] == VT_PRESSURE ? 1 : 0;
        if (i < sx - isP)
          for (long long j : bottom) layers[i + isP].push_back(j - (sx - i - isP) * dirZ);
        for (long long j : top) layers[sx + 1 + i].push_back(j + (i + 1) * dirZ);
      }
      if (i < sx - 1) {
        for (auto& d : offset) d--;
        if (typeArr[type] == VT_PRESSURE) {
          if (offset[0] < 0) {
            active.push_back(active.back() + 1);
            active.erase(active.begin());
            offset.push_back(rowLen[active.back()]);
            offset.erase(offset.begin());
          }
        } else {
          if (offset[0] < 0) {
            active.erase(active.begin());
            offset.erase(offset.begin());
          } else if (offset[0] == 0) {
            active.push_back(active.back() + 1);
            offset.push_back(rowLen[active.back()]);
          }
        }
      }
    }
  }
  nodes[0].pop_back();
  nodes[0].erase(nodes[0].begin());
  nodes[1].pop_back();
  nodes[1].erase(nodes[1].begin());
  nodes[2].pop_back();
  nodes[3].pop_back();
  nodes[3].erase(nodes[3].begin());
  template_.clear();
  template_.emplace_back();
  for (int i = 0; i < dof; ++i)
    if (variableType_[i] == VT_W) {
      for (long long v : nodes[2].front()) template_.back().push_back(v + i);
      nodes[2].erase(nodes[2].begin());
      break;
    }
  for (int j = 0; j < 2 * sx - 1; ++j) {
    template_.emplace_back();
    for (int i = 0; i < dof; ++i)
      for (int type = 0; type < 4; ++type)
        if (variableType_[i] == typeArr[type])
          for (long long v : nodes[type][j]) template_.back().push_back(v + i);
    std::sort(template_.back().begin(), template_.back().end());
  }
}

// solveGroups :567-654: classify every template node by the set of neighbouring domains it lies in
void SkewCartesianPartitioner::solveGroups() {
  const long long nx = sx_ * 4;
  const long long dirX = (long long)dof_ * sx_, dirY = (long long)dof_ * nx * sx_, dirZ = (long long)dof_ * nx * nx * sx_;
  const long long first = dirX + dirY + dirZ;
  const long long d1 = (dirY + dirX) / 2, d2 = (dirY - dirX) / 2 + dirZ, d3 = dirZ;
  const long long pos[27] = {0, -d3, d3, -d2, -d2 - d3, -d2 + d3, d2, d2 - d3, d2 + d3,
                             -d1, -d1 - d3, -d1 + d3, -d1 - d2, -d1 - d2 - d3, -d1 - d2 + d3, -d1 + d2,
                             -d1 + d2 - d3, -d1 + d2 + d3, d1, d1 - d3, d1 + d3, d1 - d2,
                             d1 - d2 - d3, d1 - d2 + d3, d1 + d2, d1 + d2 - d3, d1 + d2 + d3};
  std::vector<long long> temp;
  for (auto& layer : template_)
    for (long long v : layer) temp.push_back(v + first);
  std::vector<long long> sorted(temp);
  std::sort(sorted.begin(), sorted.end());
  std::vector<std::vector<long long>> groups(1);
  std::vector<unsigned long> domains(1, 1ul);
  for (long long node : temp) {
    unsigned long bits = 0;
    for (int i = 0; i < 27; ++i)
      if (std::binary_search(sorted.begin(), sorted.end(), node - pos[i])) bits += 1ul << i;
    bool found = false;
    for (size_t g = 0; g < groups.size(); ++g)
      if (domains[g] == bits) {
        groups[g].push_back(node);
        found = true;
        break;
      }
    if (!found) {
      groups.emplace_back(1, node);
      domains.push_back(bits);
    }
  }
  groupsT_.clear();
  groupsT_.emplace_back(1, groups[0]);
  for (size_t g = 1; g < groups.size(); ++g) {
    groupsT_.emplace_back(dof_);
    for (long long node : groups[g]) groupsT_.back()[((node % dof_) + dof_) % dof_].push_back(node);
  }
}

// GetGroups :656-812
void SkewCartesianPartitioner::getGroups(int localSd, std::vector<gidx>& interior,
                                         std::vector<SepGroup>& out) const {
  interior.clear();
  out.clear();
  const int gsd = sdMap_[localSd];
  int sdx, sdy, sdz;
  subdomainPosition(gsd, sx_, sy_, sz_, sdx, sdy, sdz);
  const long long nx = 4 * sx_;
  std::vector<std::vector<std::vector<gidx>>> groups;
  for (auto const& cat : groupsT_) {
    groups.emplace_back();
    for (auto const& group : cat) {
      groups.back().emplace_back();
      for (long long node : group) {
        const int var = (int)(node % dof_);
        int x = (int)((node / dof_) % nx) + sdx - 1 - sx_;
        int y = (int)((node / dof_ / nx) % nx) + sdy - 1 - 3 * sx_ / 2;
        int z = (int)(node / dof_ / nx / nx) + sdz - 2 * sx_;
        if (perio_ & X_PERIO) x = (x + nx_) % nx_;
        if (perio_ & Y_PERIO) y = (y + ny_) % ny_;
        if (perio_ & Z_PERIO) z = (z + nz_) % nz_;
        if (x >= 0 && x < nx_ && y >= 0 && y < ny_ && z >= 0 && z < nz_)
          groups.back().back().push_back((gidx)x * dof_ + (gidx)nx_ * y * dof_ + (gidx)nx_ * ny_ * z * dof_ + var);
      }
    }
  }
  // first pressure nodes of the interior become retained (singleton) groups.  The reference erases from
  // the vector it iterates over, so the element sliding into the erased slot is skipped: same here.
  {
    int retained = 0;
    std::vector<gidx>& in0 = groups[0][0];
    for (size_t it = 0; it < in0.size(); ++it) {
      const gidx node = in0[it];
      if (variableType_[(int)(((node % dof_) + dof_) % dof_)] == VT_PRESSURE) {
        groups.emplace_back();
        groups.back().emplace_back(1, node);
        std::vector<gidx>& in = groups[0][0];  // (emplace_back may have moved the outer vector)
        in.erase(in.begin() + it);
        if (++retained >= retainPressures_) break;
      }
    }
  }
  interior = groups[0][0];
  auto cellOwner = [&](gidx node) {
    gidx cell = node / dof_;
    return subdomainId(sx_, sy_, sz_, (int)(cell % nx_), (int)((cell / nx_) % ny_), (int)(cell / ((gidx)nx_ * ny_)));
  };
  int type = 1;
  for (size_t i = 1; i < groups.size(); ++i) {
    type++;
    for (auto const& group : groups[i]) {
      std::vector<std::pair<int, SepGroup>> parts;  // keyed by owner subdomain, kept sorted (std::map order)
      for (gidx node : group) {
        const int owner = cellOwner(node);
        auto it = std::lower_bound(parts.begin(), parts.end(), owner,
                                   [](const std::pair<int, SepGroup>& a, int b) { return a.first < b; });
        if (it != parts.end() && it->first == owner) {
          it->second.nodes.push_back(node);
        } else {
          SepGroup g;
          g.type = linkVelocities_ ? type : -1;
          g.nodes.push_back(node);
          parts.insert(it, std::make_pair(owner, g));
        }
      }
      for (auto& pr : parts) {
        if (rx_ > 1) {
          if (!linkVelocities_) type++;
          const int len = (int)pr.second.nodes.size();
          const int newLen = std::max((len + rx_ - 1) / rx_, 1);
          const int numParts = (len - 1) / newLen + 1;
          for (int j = 0; j < numParts; ++j) {
            SepGroup g2;
            g2.type = (linkVelocities_ || linkRetained_) ? type : -1;
            for (int q = j * newLen; q < (j + 1) * newLen && q < len; ++q) g2.nodes.push_back(pr.second.nodes[q]);
            out.push_back(g2);
          }
        } else {
          out.push_back(pr.second);
        }
      }
    }
  }
  // velocity nodes on the far (non-periodic) boundaries are Dirichlet rows, not separators
  for (auto& g : out) {
    std::vector<gidx> copy = g.nodes;
    for (gidx node : copy) {
      const int x = (int)((node / dof_) % nx_), y = (int)((node / dof_ / nx_) % ny_);
      const int z = (int)(node / dof_ / nx_ / ny_);
      const int vt = variableType_[(int)(node % dof_)];
      const bool hit = (dof_ > 1 && x == nx_ - 1 && vt == VT_U && !(perio_ & X_PERIO)) ||
                       (dof_ > 1 && y == ny_ - 1 && vt == VT_V && !(perio_ & Y_PERIO)) ||
                       (nz_ > 1 && dof_ > 1 && z == nz_ - 1 && vt == VT_W && !(perio_ & Z_PERIO));
      if (hit) {
        if (subdomainId(sx_, sy_, sz_, x, y, z) == gsd) interior.push_back(node);
        g.nodes.erase(std::remove(g.nodes.begin(), g.nodes.end(), node), g.nodes.end());
      }
    }
  }
  std::sort(interior.begin(), interior.end());
}

}  // namespace hymls
