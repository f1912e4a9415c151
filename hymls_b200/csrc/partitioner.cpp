#include "partitioner.hpp"

#include "hostpar.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
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

// ---------------------------------------------------------------------------------------------
// Cartesian brick arithmetic (behaviour of src/HYMLS_CartesianPartitioner.cpp:80-121): bricks of (bx,by,bz)
// cells numbered x-fastest; the last brick of an axis is short when n is not a multiple of the brick size
// ---------------------------------------------------------------------------------------------
static inline int bricksAlong(int n, int b) { return (n - 1) / b + 1; }

int CartesianPartitioner::subdomainPosition(int sd, int bx, int by, int bz, int& x, int& y, int& z) const {
  const int px = bricksAlong(nx_, bx), py = bricksAlong(ny_, by), pz = bricksAlong(nz_, bz);
  x = (sd % px) * bx;
  y = (sd / px % py) * by;
  z = (sd / px / py % pz) * bz;
  return 0;
}
int CartesianPartitioner::subdomainId(int bx, int by, int bz, int x, int y, int z) const {
  const int px = bricksAlong(nx_, bx), py = bricksAlong(ny_, by);
  return x / bx + px * (y / by + py * (z / bz));
}
int CartesianPartitioner::numGlobalParts(int bx, int by, int bz) const {
  return bricksAlong(nx_, bx) * bricksAlong(ny_, by) * bricksAlong(nz_, bz);
}
int CartesianPartitioner::pid(gidx gid) const {
  const gidx cell = gid / dof_;
  return pidMap_[subdomainId(sx_, sy_, sz_, (int)(cell % nx_), (int)(cell / nx_ % ny_), (int)(cell / nx_ / ny_ % nz_))];
}

// ---------------------------------------------------------------------------------------------
// Subdomain -> rank map.  Behaviour of BasePartitioner::CreatePIDMap (src/HYMLS_BasePartitioner.cpp:348-586),
// written as four phases over a LADDER of brick grids fine * c^k (c = smallest integer root of the coarsening
// factor), coarsest (one brick) first:
//   1. walking down the ladder, every brick is represented by its ANCHOR, the fine subdomain holding the
//      brick's origin; anchors receive ranks in order of first appearance.  The walk stops at the last grid
//      whose anchors still fit into the P ranks (the "accepted" grid);
//   2. ranks left over are dealt cyclically to the anchors in ascending subdomain order, so an accepted brick
//      may own several ranks;
//   3. the bricks of the next finer grid (the "deal" grid; the accepted one if there is none) take a rank from
//      their accepted parent brick, round-robin in brick order;
//   4. every fine subdomain inherits the rank of the deal brick it lies in.
// Works through the virtual position / id functions, so the skew partitioner shares it.
// ---------------------------------------------------------------------------------------------
static int smallestIntegerRoot(int c) {
  for (int base = 2; base < c; ++base) {
    long long pw = base;
    while (pw < c) pw *= base;
    if (pw == c) return base;
  }
  return c;
}

int CartesianPartitioner::brickAnchor(int b, int bx, int by, int bz) const {
  int x, y, z;
  subdomainPosition(b, bx, by, bz, x, y, z);
  x = ((x % nx_) + nx_) % nx_;
  y = ((y % ny_) + ny_) % ny_;
  z = ((z % nz_) + nz_) % nz_;
  return subdomainId(sx_, sy_, sz_, x, y, z);
}

void CartesianPartitioner::createPidMap() {
  const int nfine = numGlobalParts(sx_, sy_, sz_);
  const int P = nprocsComm_;
  if (P == 1 || nfine == 1) {
    nprocs_ = 1;
    pidMap_.assign(nfine, 0);
    return;
  }
  struct Grid { int bx, by, bz; };
  std::vector<Grid> ladder(1, Grid{sx_, sy_, sz_});  // ladder[0] = fine grid, back() = a single brick
  {
    const int fx = smallestIntegerRoot(cx_), fy = smallestIntegerRoot(cy_), fz = smallestIntegerRoot(cz_);
    while (ladder.back().bx < nx_ || ladder.back().by < ny_ || ladder.back().bz < nz_) {
      Grid g = ladder.back();
      g.bx *= fx;
      g.by *= fy;
      if (nz_ > 1) g.bz *= fz;
      ladder.push_back(g);
    }
  }
  auto anchorsOf = [&](const Grid& g) {
    std::vector<int> a(numGlobalParts(g.bx, g.by, g.bz));
    for (int b = 0; b < (int)a.size(); ++b) a[b] = brickAnchor(b, g.bx, g.by, g.bz);
    return a;
  };
  // phase 1
  std::vector<int> firstRank(nfine, -1);
  int nOwners = 0;
  int accepted = (int)ladder.size() - 1, deal = accepted;
  for (int lv = (int)ladder.size() - 1; lv >= 0; --lv) {
    std::vector<int> fresh;
    for (int a : anchorsOf(ladder[lv]))
      if (firstRank[a] < 0 && std::find(fresh.begin(), fresh.end(), a) == fresh.end()) fresh.push_back(a);
    deal = lv;
    if (nOwners + (int)fresh.size() > P) break;  // this grid does not fit: it becomes the deal grid
    for (int a : fresh) firstRank[a] = nOwners++;
    accepted = lv;
  }
  // phase 2
  std::vector<int> owners;
  for (int sd = 0; sd < nfine; ++sd)
    if (firstRank[sd] >= 0) owners.push_back(sd);
  std::vector<std::vector<int>> ranksOf(nfine);
  for (int sd : owners) ranksOf[sd].push_back(firstRank[sd]);
  for (int r = nOwners, turn = 0; r < P; ++r, ++turn) ranksOf[owners[turn % owners.size()]].push_back(r);
  // phase 3
  pidMap_.assign(nfine, -1);
  std::vector<int> cursor(nfine, 0);
  const Grid gd = ladder[deal], ga = ladder[accepted];
  auto wrappedOrigin = [&](int b, const Grid& g, int& x, int& y, int& z) {
    subdomainPosition(b, g.bx, g.by, g.bz, x, y, z);
    x = ((x % nx_) + nx_) % nx_;
    y = ((y % ny_) + ny_) % ny_;
    z = ((z % nz_) + nz_) % nz_;
  };
  const int nDeal = numGlobalParts(gd.bx, gd.by, gd.bz);
  for (int b = 0; b < nDeal; ++b) {
    int x, y, z;
    wrappedOrigin(b, gd, x, y, z);
    const int a = subdomainId(sx_, sy_, sz_, x, y, z);
    if (pidMap_[a] >= 0) continue;
    const int parent = brickAnchor(subdomainId(ga.bx, ga.by, ga.bz, x, y, z), ga.bx, ga.by, ga.bz);
    if (ranksOf[parent].empty()) throw argError("CreatePIDMap: a brick without an owning rank");
    pidMap_[a] = ranksOf[parent][cursor[parent]++ % ranksOf[parent].size()];
  }
  // phase 4
  for (int sd = 0; sd < nfine; ++sd) {
    if (pidMap_[sd] >= 0) continue;
    int x, y, z;
    wrappedOrigin(sd, ladder[0], x, y, z);
    int src = subdomainId(sx_, sy_, sz_, x, y, z);
    if (pidMap_[src] < 0) src = brickAnchor(subdomainId(gd.bx, gd.by, gd.bz, x, y, z), gd.bx, gd.by, gd.bz);
    if (pidMap_[src] < 0) throw argError("CreatePIDMap: a subdomain outside every brick");
    pidMap_[sd] = pidMap_[src];
  }
  std::vector<char> seen(P, 0);
  nprocs_ = 0;
  for (int r : pidMap_)
    if (!seen[r]) { seen[r] = 1; ++nprocs_; }
}

void CartesianPartitioner::partition() {
  createPidMap();
  sdMap_.clear();
  for (int sd = 0; sd < (int)pidMap_.size(); ++sd)
    if (pidMap_[sd] == mypid_) sdMap_.push_back(sd);
}

// ---------------------------------------------------------------------------------------------
// Groups of a Cartesian subdomain.  Behaviour of CartesianPartitioner::GetGroups
// (src/HYMLS_CartesianPartitioner.cpp:224-408), formulated as a tensor product: every axis of the subdomain
// is cut into PIECES
//     [plane of the previous subdomain] [inner piece 0] ... [inner piece r-1] [own separator plane]
// (no previous plane at a non-periodic near wall; at a non-periodic far wall the own plane does not exist and
// its cells join the last inner piece).  A (z-piece, y-piece, x-piece, variable) product is interior when all
// three pieces are inner (pressures: two of three), otherwise one separator group whose type encodes the
// three piece kinds in base 3.  Products are visited z-piece outermost, variable innermost: that order is the
// group order of the reference and decides which pressure nodes are retained.
// ---------------------------------------------------------------------------------------------
namespace {
struct AxisPiece {
  int kind;    // 0 previous plane, 1 inner, 2 own plane
  int lo, hi;  // cell offsets [lo, hi) from the subdomain origin
};
std::vector<AxisPiece> cutAxis(int origin, int extent, int n, int pieces, bool periodic) {
  std::vector<AxisPiece> out;
  if (periodic || origin > 0) out.push_back({0, -1, 0});
  const bool farWall = !periodic && origin + extent + 1 == n;
  const int width = std::max((extent + pieces - 1) / pieces, 1);
  for (int q = 0; q < pieces; ++q) {
    const int lo = std::min(q * width, extent);
    int hi = std::min((q + 1) * width, extent);
    if (farWall && q == pieces - 1) ++hi;
    if (hi > lo) out.push_back({1, lo, hi});
  }
  if (!farWall) out.push_back({2, extent, extent + 1});
  return out;
}
}  // namespace

void CartesianPartitioner::getGroups(int localSd, std::vector<gidx>& interior,
                                     std::vector<SepGroup>& groups) const {
  interior.clear();
  groups.clear();
  int ox, oy, oz;
  subdomainPosition(sdMap_[localSd], sx_, sy_, sz_, ox, oy, oz);
  const int ex = std::min(nx_ - ox, sx_) - 1, ey = std::min(ny_ - oy, sy_) - 1, ez = std::min(nz_ - oz, sz_) - 1;
  if (ex == 0 || ey == 0 || (ez == 0 && nz_ > 1)) throw argError("Can't have a subdomain of size 1");
  const std::vector<AxisPiece> cutX = cutAxis(ox, ex, nx_, std::max(rx_, 1), perio_ & X_PERIO);
  const std::vector<AxisPiece> cutY = cutAxis(oy, ey, ny_, std::max(ry_, 1), perio_ & Y_PERIO);
  const std::vector<AxisPiece> cutZ = cutAxis(oz, ez, nz_, std::max(rz_, 1), perio_ & Z_PERIO);
  // grid coordinates of the offsets -1 .. extent (periodic wrap)
  auto coords = [](int origin, int extent, int n) {
    std::vector<gidx> c(extent + 2);
    for (int o = -1; o <= extent; ++o) c[o + 1] = (gidx)((origin + o + n) % n);
    return c;
  };
  const std::vector<gidx> X = coords(ox, ex, nx_), Y = coords(oy, ey, ny_), Z = coords(oz, ez, nz_);
  std::vector<gidx> retained;
  for (const AxisPiece& pz : cutZ)
    for (const AxisPiece& py : cutY)
      for (const AxisPiece& px : cutX) {
        const bool touchesPrev = px.kind == 0 || py.kind == 0 || pz.kind == 0;
        const int innerAxes = (px.kind == 1) + (py.kind == 1) + (pz.kind == 1);
        for (int d = 0; d < dof_; ++d) {
          const int vt = variableType_[d];
          const bool cellCentred = vt == VT_PRESSURE || vt == VT_INTERIOR;
          if (cellCentred && touchesPrev) continue;  // those planes carry the neighbour's cell-centred unknowns
          const bool isInterior = innerAxes == 3 || vt == VT_INTERIOR ||
                                  (vt == VT_PRESSURE && (innerAxes == 2 || retainPressures_ > 1));
          int target = -1, targetOdd = -1;  // group indices (B-grid: odd cells go to a second group)
          if (!isInterior) {
            int type = linkRetained_ ? 2 * dof_ * (px.kind + 3 * (py.kind + 3 * pz.kind)) : -1000;
            const bool velocity = vt == VT_U || vt == VT_V || vt == VT_W;
            const bool shared = (linkVelocities_ && velocity) || (linkTubePressures_ && vt == VT_PRESSURE);
            if (!shared) type += 2 * d;
            target = (int)groups.size();
            groups.emplace_back();
            groups.back().type = type;
            if (bgrid_) {
              targetOdd = (int)groups.size();
              groups.emplace_back();
              groups.back().type = type + 1;
            }
          }
          for (int k = pz.lo; k < pz.hi; ++k)
            for (int j = py.lo; j < py.hi; ++j)
              for (int i = px.lo; i < px.hi; ++i) {
                const gidx gid = d + dof_ * (X[i + 1] + nx_ * (Y[j + 1] + (gidx)ny_ * Z[k + 1]));
                if (vt == VT_PRESSURE && i >= 0 && j >= 0 && k >= 0 && (int)retained.size() < retainPressures_)
                  retained.push_back(gid);
                else if (targetOdd >= 0 && (i + ox + j + oy) % 2)
                  groups[targetOdd].nodes.push_back(gid);
                else if (target < 0)
                  interior.push_back(gid);
                else
                  groups[target].nodes.push_back(gid);
              }
        }
      }
  size_t kept = 0;
  for (size_t g = 0; g < groups.size(); ++g)
    if (!groups[g].nodes.empty()) {
      if (kept != g) std::swap(groups[kept], groups[g]);
      ++kept;
    }
  groups.resize(kept);
  for (gidx g : retained) {  // retained pressures: singleton groups that are never linked
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
  // the per-subdomain group lists are independent (GetGroups is const): build and sort them in parallel, then
  // merge sequentially in subdomain order (the order decides which subdomain owns a shared group)
  std::vector<std::vector<gidx>> allInterior(nsd);
  std::vector<std::vector<SepGroup>> allGroups(nsd);
  std::vector<std::string> errors(64);
  const auto tg0 = std::chrono::steady_clock::now();
  parallelFor(nsd, [&](int64_t s0, int64_t s1, int t) {
    try {
      for (int64_t sd = s0; sd < s1; ++sd) {
        part.getGroups((int)sd, allInterior[sd], allGroups[sd]);
        if (filter) {  // coarser levels keep ~1 % of the grid nodes: drop the others here, in parallel
          auto gone = [&](gidx g) { return !present[g]; };
          std::vector<gidx>& in = allInterior[sd];
          in.erase(std::remove_if(in.begin(), in.end(), gone), in.end());
          for (auto& grp : allGroups[sd])
            grp.nodes.erase(std::remove_if(grp.nodes.begin(), grp.nodes.end(), gone), grp.nodes.end());
        }
        std::sort(allInterior[sd].begin(), allInterior[sd].end());
        for (auto& grp : allGroups[sd]) std::sort(grp.nodes.begin(), grp.nodes.end());
      }
    } catch (const std::exception& e) {
      errors[t & 63] = e.what();
    }
  }, 16);
  for (const std::string& e : errors)
    if (!e.empty()) throw Error(HYMLS_B200_ERR_ARG, e);
  if (getenv("HYMLS_B200_VERBOSE_SYM"))
    fprintf(stderr, "[hymls_b200 sym] getGroups (parallel)                    %.3f s\n",
            std::chrono::duration<double>(std::chrono::steady_clock::now() - tg0).count());
  // Merge.  Which subdomain owns a shared group is decided in subdomain order, but only the GROUP bookkeeping is
  // sequential (~10^5 groups); the node lists (~10^7 GIDs) are copied in parallel once every offset is known.
  // (the parallel phase above already removed the nodes that do not exist on this level)
  gidx maxGid = -1;
  {
    std::vector<gidx> tmax(64, -1);
    parallelFor(nsd, [&](int64_t s0, int64_t s1, int t) {
      gidx m = -1;
      for (int64_t sd = s0; sd < s1; ++sd)
        for (const auto& grp : allGroups[sd])
          if (!grp.nodes.empty()) m = std::max(m, grp.nodes.front());
      tmax[t & 63] = std::max(tmax[t & 63], m);
    }, 16);
    for (gidx m : tmax) maxGid = std::max(maxGid, m);
  }
  std::vector<int> uniqueByFirst((size_t)(maxGid + 1), -1);  // first GID of a group -> unique id
  std::vector<int> uniqSrcSd, uniqSrcGrp;                     // where the node list of a unique group comes from
  std::vector<int64_t> ovlPtr(nsd + 1, 0);
  std::vector<int64_t> sdFirstUniq(nsd + 1, 0);
  H.intPtr.resize(nsd + 1);
  H.sdGrpPtr.resize(nsd + 1);
  for (int sd = 0; sd < nsd; ++sd) {
    H.intPtr[sd + 1] = H.intPtr[sd] + (int64_t)allInterior[sd].size();
    int64_t newSep = 0;
    const std::vector<SepGroup>& groups = allGroups[sd];
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      const SepGroup& grp = groups[gi];
      if (grp.nodes.empty()) continue;  // empty after filtering: removed (:241-244)
      const int64_t len = (int64_t)grp.nodes.size();
      H.grpPtr.push_back(H.grpPtr.back() + len);
      H.grpType.push_back(grp.type);
      int& slot = uniqueByFirst[grp.nodes.front()];
      if (slot < 0) {
        slot = (int)H.uniqOwnerSd.size();
        H.uniqOwnerSd.push_back(sd);
        H.uniqType.push_back(grp.type);
        H.uniqPtr.push_back(H.uniqPtr.back() + len);
        uniqSrcSd.push_back(sd);
        uniqSrcGrp.push_back((int)gi);
        newSep += len;
      } else if (H.uniqPtr[slot + 1] - H.uniqPtr[slot] != len) {
        // the reference identifies groups by their first GID only (:261-271); a mismatch in the
        // node list would make its maps inconsistent, so we refuse it.
        throw Error(HYMLS_B200_ERR_ARG, "separator group seen with two different node lists");
      }
      H.grpUnique.push_back(slot);
    }
    H.sdGrpPtr[sd + 1] = (int64_t)H.grpType.size();
    sdFirstUniq[sd + 1] = (int64_t)H.uniqOwnerSd.size();
    ovlPtr[sd + 1] = ovlPtr[sd] + (int64_t)allInterior[sd].size() + newSep;
  }
  H.intGid.resize(H.intPtr[nsd]);
  H.grpGid.resize(H.grpPtr.back());
  H.sepGid.resize(H.uniqPtr.back());
  H.overlappingGid.resize(ovlPtr[nsd]);
  parallelFor(nsd, [&](int64_t s0, int64_t s1, int) {
    for (int64_t sd = s0; sd < s1; ++sd) {
      std::vector<gidx>& interior = allInterior[sd];
      std::copy(interior.begin(), interior.end(), H.intGid.begin() + H.intPtr[sd]);
      gidx* ovl = H.overlappingGid.data() + ovlPtr[sd];
      ovl = std::copy(interior.begin(), interior.end(), ovl);
      int64_t g = H.sdGrpPtr[sd];
      for (const SepGroup& grp : allGroups[sd]) {
        if (grp.nodes.empty()) continue;
        std::copy(grp.nodes.begin(), grp.nodes.end(), H.grpGid.begin() + H.grpPtr[g]);
        ++g;
      }
      for (int64_t u = sdFirstUniq[sd]; u < sdFirstUniq[sd + 1]; ++u) {  // the unique groups this subdomain owns
        const std::vector<gidx>& nodes = allGroups[uniqSrcSd[u]][uniqSrcGrp[u]].nodes;
        std::copy(nodes.begin(), nodes.end(), H.sepGid.begin() + H.uniqPtr[u]);
        ovl = std::copy(nodes.begin(), nodes.end(), ovl);
      }
    }
  }, 16);
  parallelFor(nsd, [&](int64_t s0, int64_t s1, int) {  // release the per-subdomain lists (in parallel: ~10^6 frees)
    for (int64_t sd = s0; sd < s1; ++sd) {
      std::vector<gidx>().swap(allInterior[sd]);
      std::vector<SepGroup>().swap(allGroups[sd]);
    }
  }, 16);
}

}  // namespace hymls

// =============================================================================================
// Skew Cartesian partitioner
// =============================================================================================
namespace hymls {

CartesianPartitioner* makePartitioner(ParameterList& params, int level, int nprocs, int mypid) {
  std::string method = params.sublist("Preconditioner").get("Partitioner", "Cartesian");
  if (method == "Cartesian") return new CartesianPartitioner(params, level, nprocs, mypid);
  if (method == "Skew Cartesian") return new SkewCartesianPartitioner(params, level, nprocs, mypid);
  throw Error(HYMLS_B200_ERR_ARG, "Up to now we only support Cartesian partitioning");
}

// ---------------------------------------------------------------------------------------------
// Geometry.  In the sheared coordinates  s = x + y,  d = x - y,  t = x - y + z  the skew subdomains of size b
// are the cubes  S*b - 1 <= s < (S+1)*b - 1,  D*b <= d < (D+1)*b,  T*b <= t < (T+1)*b  (the cells a subdomain
// owns = the pressure nodes of its template).  The reference numbers them layer by layer (Z = T - D), each
// layer as interleaved rows: "even" rows of npx subdomains at (col*b, row*b) followed by "odd" rows of npx + 1
// subdomains at ((col - 1/2) b, (row + 1/2) b).  Checked against the closed forms the reference's unit tests
// pin (tests/test_oracle_skew.py) and cell by cell against the oracle (tests/test_host_maps.py).
// ---------------------------------------------------------------------------------------------
static inline int floorDiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

int SkewCartesianPartitioner::numGlobalParts(int bx, int by, int bz) const {
  const int npx = nx_ / bx, npy = ny_ / by, npz = nz_ / bz;
  const int perLayer = 2 * npx * npy + npx + npy;
  return std::max(nz_ > 1 ? perLayer * (npz + 1) : perLayer, 1);
}

int SkewCartesianPartitioner::subdomainPosition(int sd, int bx, int by, int bz, int& x, int& y, int& z) const {
  (void)bz;
  const int npx = nx_ / bx, npy = ny_ / by;
  const int perLayer = 2 * npx * npy + npx + npy, perRowPair = 2 * npx + 1;
  const int layer = perLayer > 0 ? sd / perLayer : 0;
  const int inLayer = sd - layer * perLayer;
  const int pairIdx = inLayer / perRowPair, inPair = inLayer % perRowPair;
  // a "row pair" is an odd row (npx + 1 subdomains, half a brick lower) stored AFTER the even row above it:
  // inPair < npx: even row at y = pairIdx*b - b/2 ... written in half-brick units below
  const int h = bx / 2;
  if (inPair < npx) {
    x = 2 * inPair * h;
    y = (2 * pairIdx - 1) * h + h;
  } else {
    x = (2 * (inPair - npx) - 1) * h;
    y = 2 * pairIdx * h + h;
  }
  z = layer * bx;
  // periodic images: the last odd column / last row pair / top layer coincide with the first ones
  if ((perio_ & X_PERIO) && x == nx_ - h) return 1;
  if ((perio_ & Y_PERIO) && y == ny_) return 1;
  if ((perio_ & Z_PERIO) && z == nz_) return 1;
  return 0;
}

int SkewCartesianPartitioner::subdomainId(int bx, int by, int bz, int x, int y, int z) const {
  const int b = bx;
  const int npx = nx_ / bx, npy = ny_ / by, npz = nz_ / bz;
  const int perLayer = 2 * npx * npy + npx + npy, perRowPair = 2 * npx + 1;
  const int S = floorDiv(x + y + 1, b), D = floorDiv(x - y, b), T = floorDiv(x - y + z, b);
  int layer = T - D;
  // cube (S, D) of the s-d plane: S + D even -> even row at (col, row) = ((S+D)/2, (S-D)/2),
  //                               S + D odd  -> odd row, col = (S+D+1)/2 in 0..npx, below even row (S-D+1)/2
  int pairIdx, inPair;
  if (((S + D) & 1) == 0) {
    pairIdx = (S - D) / 2;
    inPair = (S + D) / 2;
    if ((perio_ & Y_PERIO) && pairIdx == npy) pairIdx = 0;
  } else {
    pairIdx = (S - D - 1) / 2;
    int col = (S + D + 1) / 2;
    if ((perio_ & X_PERIO) && col == npx) col = 0;
    inPair = npx + col;
  }
  if ((perio_ & Z_PERIO) && layer == npz) layer = 0;
  return layer * perLayer + pairIdx * perRowPair + inPair;
}

// ---------------------------------------------------------------------------------------------
// Template of a subdomain = the nodes it touches (its interior and all separators around it), relative to the
// subdomain position.  Per variable type the reference's template is exactly the set of lattice points of a
// polytope in (s, d, t, z) (b = sx; found by fitting the four plane families to the reference's node lists
// for sx = 4, 6, 8 and asserted equal to the oracle's lists in tests/test_oracle_skew.py):
//     P:  -1 <= s <= b-2    0 <= d <= b-1    0 <= t <= b-1    |z| <= b-1        (the owned cells)
//     U:  -2 <= s <= b-2   -1 <= d <= b-1   -1 <= t <= b-1    |z| <= b-1        (+ west faces)
//     V:  -2 <= s <= b-2    0 <= d <= b      0 <= t <= b      |z| <= b-1        (+ south faces)
//     W:  -2 <= s <= b-1   -1 <= d <= b     -1 <= t <= b-1   -b <= z <= b-1     (+ bottom faces), without the
//         nodes of even t on the four side planes s = -2, s = b-1, d = -1, d = b
// In 2D only z = 0 exists.
// ---------------------------------------------------------------------------------------------
namespace {
struct SkewShape {
  int b;
  bool flat;  // nz == 1
  bool contains(int vt, int x, int y, int z) const {
    if (flat && z != 0) return false;
    const int s = x + y, d = x - y, t = d + z;
    switch (vt) {
      case VT_PRESSURE: return s >= -1 && s <= b - 2 && d >= 0 && d <= b - 1 && t >= 0 && t <= b - 1 && z > -b && z < b;
      case VT_U: return s >= -2 && s <= b - 2 && d >= -1 && d <= b - 1 && t >= -1 && t <= b - 1 && z > -b && z < b;
      case VT_V: return s >= -2 && s <= b - 2 && d >= 0 && d <= b && t >= 0 && t <= b && z > -b && z < b;
      case VT_W:
        if (!(s >= -2 && s <= b - 1 && d >= -1 && d <= b && t >= -1 && t <= b - 1 && z >= -b && z < b)) return false;
        return (t & 1) || !(s == -2 || s == b - 1 || d == -1 || d == b);
      default: return false;
    }
  }
};
}  // namespace

// Classifies every template node by the set of neighbouring subdomains (lattice translates a*e1 + b*e2 + c*e3,
// a, b, c in {-1, 0, 1}, e1 = (h, h, 0), e2 = (-h, h, 2h), e3 = (0, 0, 2h), h = sx/2) whose templates contain it
// as well.  Class 0 = no neighbour (interior); the other classes are numbered in order of first appearance
// when the template is scanned by ascending (z, y, x, variable) -- the reference's group order.
void SkewCartesianPartitioner::classifyTemplate() {
  const SkewShape shape{sx_, nz_ <= 1};
  const int b = sx_, h = sx_ / 2;
  struct Shift { int x, y, z; };
  std::vector<Shift> shifts;
  for (int a = -1; a <= 1; ++a)
    for (int bb = -1; bb <= 1; ++bb)
      for (int c = -1; c <= 1; ++c)
        if (a || bb || c) shifts.push_back({a * h - bb * h, a * h + bb * h, bb * b + c * b});
  std::vector<uint32_t> classMask;  // classMask[k-1] = neighbour set of class k
  std::vector<TemplateNode> scan;
  for (int z = -b; z < b; ++z)
    for (int y = -2 * b; y <= 2 * b; ++y)
      for (int x = -2 * b; x <= 2 * b; ++x)
        for (int v = 0; v < dof_; ++v) {
          const int vt = variableType_[v];
          if (!shape.contains(vt, x, y, z)) continue;
          uint32_t mask = 0;
          for (size_t q = 0; q < shifts.size(); ++q)
            if (shape.contains(vt, x - shifts[q].x, y - shifts[q].y, z - shifts[q].z)) mask |= 1u << q;
          int cls = 0;
          if (mask) {
            size_t k = 0;
            while (k < classMask.size() && classMask[k] != mask) ++k;
            if (k == classMask.size()) classMask.push_back(mask);
            cls = (int)k + 1;
          }
          scan.push_back({x, y, z, v, cls});
        }
  ncls_ = (int)classMask.size() + 1;
  // bucket by (class, variable), keeping the scan order inside a bucket
  clsVarPtr_.assign((size_t)ncls_ * dof_ + 1, 0);
  for (const TemplateNode& t : scan) clsVarPtr_[(size_t)t.cls * dof_ + t.var + 1]++;
  for (size_t q = 1; q < clsVarPtr_.size(); ++q) clsVarPtr_[q] += clsVarPtr_[q - 1];
  std::vector<int64_t> fill(clsVarPtr_.begin(), clsVarPtr_.end() - 1);
  tmpl_.resize(scan.size());
  for (const TemplateNode& t : scan) tmpl_[fill[(size_t)t.cls * dof_ + t.var]++] = t;
  innerScan_.clear();  // class 0 once more, all variables in one scan order (z, y, x, variable)
  for (const TemplateNode& t : scan)
    if (t.cls == 0) innerScan_.push_back(t);
}

void SkewCartesianPartitioner::partition() {
  if (sx_ != sy_ || (nz_ > 1 && sx_ != sz_)) throw argError("sx, sy and sz should be the same");
  if (sx_ % 2) throw argError("sx should be even");
  if (nx_ % sx_ || ny_ % sy_ || nz_ % sz_)
    throw argError("You are trying to partition a domain whose size is not a multiple of the subdomain size");
  createPidMap();
  sdMap_.clear();
  const int nparts = numGlobalParts(sx_, sy_, sz_);
  for (int sd = 0; sd < nparts; ++sd) {
    int x, y, z;
    if (subdomainPosition(sd, sx_, sy_, sz_, x, y, z) == 1) continue;  // periodic image of another subdomain
    if (pidMap_[brickAnchor(sd, sx_, sy_, sz_)] == mypid_) sdMap_.push_back(sd);
  }
  classifyTemplate();
}

// Groups of one subdomain (behaviour of SkewCartesianPartitioner::GetGroups, src/HYMLS_SkewCartesianPartitioner.cpp:
// 656-812): the classified template translated to the subdomain position and clipped to the grid.  Class 0 is
// the interior (its first pressure nodes become retained singleton groups); every other (class, variable)
// bucket is split by the subdomain that owns the node's cell, ascending, optionally subdivided into rx pieces.
// Velocities on a non-periodic far wall are Dirichlet rows: dropped from the groups, and kept as interior by
// the subdomain owning the cell.
void SkewCartesianPartitioner::getGroups(int localSd, std::vector<gidx>& interior,
                                         std::vector<SepGroup>& out) const {
  interior.clear();
  out.clear();
  const int me = sdMap_[localSd];
  int px, py, pz;
  subdomainPosition(me, sx_, sy_, sz_, px, py, pz);
  // translated node: returns false when it falls outside the grid
  auto place = [&](const TemplateNode& t, int& x, int& y, int& z) {
    x = t.dx + px;
    y = t.dy + py;
    z = t.dz + pz;
    if (perio_ & X_PERIO) x = (x + nx_) % nx_;
    if (perio_ & Y_PERIO) y = (y + ny_) % ny_;
    if (perio_ & Z_PERIO) z = (z + nz_) % nz_;
    return x >= 0 && x < nx_ && y >= 0 && y < ny_ && z >= 0 && z < nz_;
  };
  auto gidOf = [&](int x, int y, int z, int v) { return v + dof_ * ((gidx)x + nx_ * ((gidx)y + (gidx)ny_ * z)); };
  auto onFarWall = [&](int x, int y, int z, int vt) {
    if (dof_ <= 1) return false;
    return (vt == VT_U && x == nx_ - 1 && !(perio_ & X_PERIO)) || (vt == VT_V && y == ny_ - 1 && !(perio_ & Y_PERIO)) ||
           (vt == VT_W && nz_ > 1 && z == nz_ - 1 && !(perio_ & Z_PERIO));
  };
  // --- class 0: interior, in scan order over all variables
  struct Placed { gidx gid; bool pressure; };
  static thread_local std::vector<Placed> inner;
  inner.clear();
  for (const TemplateNode& t : innerScan_) {
    int x, y, z;
    if (place(t, x, y, z)) inner.push_back({gidOf(x, y, z, t.var), variableType_[t.var] == VT_PRESSURE});
  }
  std::vector<gidx> retained;
  {
    // The first pressure nodes of the interior are retained.  Reference quirk kept on purpose: it erases from
    // the list it iterates over, so the element right after a retained node is not examined.
    size_t q = 0;
    interior.reserve(inner.size() + 64);
    for (; q < inner.size() && (int)retained.size() < retainPressures_; ++q) {
      if (inner[q].pressure) {
        retained.push_back(inner[q].gid);
        if (++q < inner.size()) interior.push_back(inner[q].gid);
      } else {
        interior.push_back(inner[q].gid);
      }
    }
    for (; q < inner.size(); ++q) interior.push_back(inner[q].gid);
  }
  // --- separator classes.  Far-wall velocities take part in the splitting (they count towards the length of a
  // part and keep its type number alive) and are removed from the emitted groups afterwards, like the reference.
  int type = 1;
  struct Part { int owner; std::vector<gidx> nodes; std::vector<char> wall; };
  static thread_local std::vector<Part> parts;  // pool: the node vectors keep their capacity between calls
  size_t nparts = 0;
  auto emitParts = [&]() {
    if (nparts > 1)
      std::stable_sort(parts.begin(), parts.begin() + nparts, [](const Part& a, const Part& c) { return a.owner < c.owner; });
    for (size_t pi = 0; pi < nparts; ++pi) {
      Part& p = parts[pi];
      const int len = (int)p.nodes.size();
      int piece = len, groupType = linkVelocities_ ? type : -1;
      if (rx_ > 1) {
        if (!linkVelocities_) ++type;
        piece = std::max((len + rx_ - 1) / rx_, 1);
        groupType = (linkVelocities_ || linkRetained_) ? type : -1;
      }
      for (int lo = 0; lo < len; lo += piece) {
        out.emplace_back();  // (may stay empty: the caller drops empty groups)
        out.back().type = groupType;
        const int hi = std::min(len, lo + piece);
        out.back().nodes.reserve(hi - lo);
        for (int q = lo; q < hi; ++q)
          if (!p.wall[q]) out.back().nodes.push_back(p.nodes[q]);
      }
    }
    nparts = 0;
  };
  auto newPart = [&](int owner) -> Part& {
    if (nparts == parts.size()) parts.emplace_back();
    Part& p = parts[nparts++];
    p.owner = owner;
    p.nodes.clear();
    p.wall.clear();
    return p;
  };
  for (int cls = 1; cls < ncls_; ++cls) {
    ++type;
    for (int v = 0; v < dof_; ++v) {
      const int vt = variableType_[v];
      for (int64_t q = clsVarPtr_[(size_t)cls * dof_ + v]; q < clsVarPtr_[(size_t)cls * dof_ + v + 1]; ++q) {
        int x, y, z;
        if (!place(tmpl_[q], x, y, z)) continue;
        const int owner = subdomainId(sx_, sy_, sz_, x, y, z);
        const bool wall = onFarWall(x, y, z, vt);
        if (wall && owner == me) interior.push_back(gidOf(x, y, z, v));
        size_t k = 0;
        while (k < nparts && parts[k].owner != owner) ++k;
        if (k == nparts) newPart(owner);
        parts[k].nodes.push_back(gidOf(x, y, z, v));
        parts[k].wall.push_back(wall ? 1 : 0);
      }
      emitParts();
    }
  }
  for (gidx g : retained) {  // retained pressures: one more class each, through the same splitting rules
    ++type;
    Part& p = newPart(me);
    p.nodes.push_back(g);
    p.wall.push_back(0);
    emitParts();
  }
  std::sort(interior.begin(), interior.end());
}

}  // namespace hymls
