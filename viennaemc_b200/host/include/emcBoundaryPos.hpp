// Faces of the simulation box.  Interface mirrored: reference include/emcBoundaryPos.hpp.
#ifndef EMC_BOUNDARY_POS_HPP
#define EMC_BOUNDARY_POS_HPP

#include <emcUtil.hpp>

// value / 2 = dimension normal to the face, value % 2 = 0 at the origin side
enum struct emcBoundaryPos : SizeType { XMIN = 0, XMAX, YMIN, YMAX, ZMIN, ZMAX, INVALID };

#endif
