// orc_convex.h — shared declarations of the oracle's narrowphase (orc_collision.cpp, orc_convex.cpp).
// TEST INFRASTRUCTURE ONLY (see orc_math.h header).
#pragma once
#include "b2mj.h"

namespace orc {

struct Con {
  double dist, pos[3], frame[9];
};

// a convex geom as the support mapping sees it: world pose, primitive sizes, hull vertices for meshes (geom frame)
struct ConvexGeom {
  int type;
  const double *pos, *mat, *size;
  const double* vert;
  int nvert;
};

int planeCylinder(Con* con, double margin, const double* pos1, const double* mat1, const double* pos2, const double* mat2,
                  const double* size2);
int planeConvex(Con* con, double margin, const double* pos1, const double* mat1, const ConvexGeom& g);
int convexConvex(Con* con, double margin, const ConvexGeom& g1, const ConvexGeom& g2, int mpr_iterations, double mpr_tolerance);
void convexSupport(const ConvexGeom& g, const double* dir, double* res);
int hfieldConvex(Con* con, int maxcon, double margin, const double* pos1, const double* mat1, const double* hsize, int nrow,
                 int ncol, const double* data, const ConvexGeom& g2, double rbound2, int mpr_iterations, double mpr_tolerance);

}  // namespace orc
