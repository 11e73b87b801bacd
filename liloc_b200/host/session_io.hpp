// liloc_b200/host/session_io.hpp -- the on-disk session format of the reference, dependency-free.
//
// A saved session (src/mapOptmization.cpp:420-635, read back by dataManager::Session,
// include/dataManager/dataLoader.hpp:207-304) is a directory with
//     singlesession_posegraph.g2o     VERTEX_SE3:QUAT id x y z qx qy qz qw   /   EDGE_SE3:QUAT i j x y z qx qy qz qw
//     PCDs/<k>.pcd                    key cloud k, body frame (pcl::io::savePCDFileBinary of pcl::PointXYZI)
//     globalMap.pcd                   (visualisation only; not needed by the path)
// The reference reads these with PCL and GTSAM; here: a PCD reader for the `ascii`, `binary` and `binary_compressed`
// (LZF, field-major) encodings with
// x / y / z / intensity float fields in any order (padding and extra fields are skipped), a g2o vertex parser, and
// the matching writers.
#pragma once
#include <string>
#include <vector>

#include "../../include/liloc_gpu.h"
#include "map_optimization.hpp"

namespace liloc_host {

// pcl::io::loadPCDFile<pcl::PointXYZI>.  Returns false and sets err on failure.
bool loadPCD(const std::string& path, std::vector<liloc_point>& out, std::string& err);
// pcl::io::savePCDFileBinary<pcl::PointXYZI> (fields x y z intensity, 16 bytes per point, DATA binary)
bool savePCDBinary(const std::string& path, const liloc_point* pts, int n, std::string& err);

// pcl::io::savePCDFileBinary<PointTypePose> (transformations.pcd, :438 / :564): PointXYZIRPYT as PCL lays it out -- 48 bytes
// per pose {x y z pad4 intensity roll pitch yaw time(double) pad8}, the padding declared as `_` fields (include/utility.h:87-103)
bool savePosesPCDBinary(const std::string& path, const std::vector<PointTypePose>& keyPoses, std::string& err);

// Session::loadSessionGraph + initKeyPoses (dataLoader.hpp:207-268): the vertices in ascending id order as key poses
// (roll / pitch / yaw from the quaternion like gtsam::Rot3::roll/pitch/yaw, intensity = node id).
bool loadSessionGraph(const std::string& g2o_path, std::vector<PointTypePose>& keyPoses, std::string& err);
// the g2o written by save_session (:566-588): vertices, then edges between consecutive key poses, std::to_string numbers
bool saveSessionGraph(const std::string& g2o_path, const std::vector<PointTypePose>& keyPoses, std::string& err);

// Session::loadKeyCloud (dataLoader.hpp:283-304): every PCDs/*.pcd, sorted by the integer in the file name
bool loadKeyClouds(const std::string& pcd_dir, std::vector<std::vector<liloc_point>>& clouds, std::string& err);

}  // namespace liloc_host
