// Host-side direction / pattern tables for the diffuse sweep.
//
// Everything here is O(directions x layers) trigonometry that decides branches (which face a characteristic
// leaves through).  It runs once per solve on the host with glibc libm -- the library gfortran links the
// reference against -- so the branch decisions are the reference's; the device only consumes the tables.
// Replaces: equiSources.f90:1385-1553 (direction loop set-up), :2118-2361 (pix2ang_nest, rotateAngles, getAngle),
// transportRoutinesModule.f90:7-85 (setPattern), :121-218 (setRaysRefined), rotateIndicesModule.f90:7-113.
#pragma once
#include <cstdint>
#include <vector>

namespace rtb {

// Single-precision literals of the reference, widened (definitionsModule.f90:8-10).
constexpr double kPi = (double)3.141592654f;
constexpr double kHalfPi = 0.5 * kPi;
constexpr double kTwoPi = 2.0 * kPi;

// How a rotated index triple (i = sweep axis, j, k) maps to the physical (icell, jcell, kcell):
// physical component c takes rotated index src[c] (0 = i, 1 = j, 2 = k), reflected to n+1-index if refl[c].
struct ZoneMap {
  int8_t src[3];
  int8_t refl[3];
};
ZoneMap zone_map(int izone);  // izone 1..24 (rotateIndicesModule.f90:14-111)

// signed leaf-index strides of the rotated axes on a uniform n^3 grid (leaf = ((ic-1)*n + jc-1)*n + kc-1)
struct ZoneStrides {
  int64_t origin;     // leaf index of rotated (1,1,1)
  int64_t stride[3];  // per unit step of rotated i, j, k
};
ZoneStrides zone_strides(int izone, int n);
// same for an array laid out with element strides phys[0..2] along the physical x, y, z axes
ZoneStrides zone_strides_layout(int izone, int n, const int64_t phys[3]);

struct Direction {
  int64_t iray;
  int izone;
  double phi, theta;  // local angles inside the zone
  int status;
};
int healpix_center(int nside, int64_t ipix, double* phi, double* theta);  // includes the fixed 0.111/0.222 rotation
Direction classify_direction(int nAngularLevel, int64_t iray);

// One layer (or sub-layer) pattern: the 1..3 segments every cell of the layer carries.
struct RayPattern {
  double xy_x0, xy_y0, xy_len;
  double xz_x0, xz_z0, xz_len;
  double yz_y0, yz_z0, yz_len;
  int8_t xzActive, yzActive;
  int8_t xyTop, xzTop, yzTop;  // 1 = xy ray, 2 = yz ray, 3 = xz ray leaves through that face; 0 = none
  int8_t status;
};

// Patterns of all `count` layers of refinement level `level` for one direction; level 0 has n layers and
// level L has n*2^L.  `parent` = the level-1 table (nullptr for level 0).
void layer_patterns_level0(double phi, double theta, int n, std::vector<RayPattern>& out);
void layer_patterns_refine(double phi, double theta, const std::vector<RayPattern>& parent,
                           std::vector<RayPattern>& out);

}  // namespace rtb
