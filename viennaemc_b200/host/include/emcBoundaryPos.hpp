// Faces of the simulation box.  Interface mirrored: reference include/emcBoundaryPos.hpp.
#ifndef EMC_BOUNDARY_POS_HPP
#define EMC_BOUNDARY_POS_HPP

enum class emcBoundaryPos : unsigned { XMIN = 0, XMAX = 1, YMIN = 2, YMAX = 3, ZMIN = 4, ZMAX = 5 };

#endif
