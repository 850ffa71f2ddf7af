// mesh.h — mesh assets of the model compiler (library-internal).
#pragma once
#include <string>
#include <vector>

namespace b2mj {

struct MeshData {
  std::vector<double> vert;  // convex-hull vertices in the mesh frame (origin = centre of mass, axes = principal axes)
  double pos[3] = {0, 0, 0}, quat[4] = {1, 0, 0, 0};  // mesh frame in the coordinates of the asset
  double volume = 0;
  double inertia[3] = {0, 0, 0};  // principal moments for unit density
  double aabb[3] = {0, 0, 0};     // half sizes of the bounding box in the mesh frame (geom_size of a mesh geom)
  double rbound = 0;
};

bool convex_hull(const std::vector<double>& pts, std::vector<int>& tri, std::string& err);
bool mesh_process(const std::vector<double>& raw, const double scale[3], MeshData& out, std::string& err);
bool mesh_read_file(const std::string& path, std::vector<double>& pts, std::string& err);

// directory mesh files are resolved against: set by b2mj_model_from_xml_file for the duration of a compile
void set_model_dir(const std::string& dir);
const std::string& model_dir();

}  // namespace b2mj
