// liloc_b200/host/param_server.hpp -- ROS-free ParamServer with the reference's parameter names,
// defaults and units (include/utility.h:117-304).  The reference reads them from the ROS parameter
// server, which the launch files fill from config/*.yaml; here the same yaml files are parsed directly
// (two-level maps, scalars, inline lists).
#pragma once
#include <cfloat>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace liloc_host {

enum class SensorType { VELODYNE, OUSTER, LIVOX, ROBOSENSE, MULRAN };   // include/utility.h:108
enum class ModeType { LIO, RELO };                                      // include/utility.h:111

class ParamServer {
public:
    // System/*
    std::string pointCloudTopic = "points_raw", imuTopic = "imu_correct", odomTopic = "odometry/imu";
    ModeType mode = ModeType::LIO;
    int numberOfCores = 4;
    std::string lidarFrame = "base_link", baselinkFrame = "base_link", odometryFrame = "odom", mapFrame = "map";
    std::string savePCDDirectory = "/Downloads/LOAM/", saveSessionDirectory = "/Downloads/LOAM/";
    // Sensors/*
    SensorType sensor = SensorType::VELODYNE;
    int N_SCAN = 16, Horizon_SCAN = 1800;
    bool have_ring_time_channel = true;
    int downsampleRate = 1, point_filter_num = 3;
    float lidarMinRange = 1.0f, lidarMaxRange = 1000.0f;
    int imuType = 0;
    float imuRate = 500.0f, imuAccNoise = 0.01f, imuGyrNoise = 0.001f, imuAccBiasN = 0.0002f, imuGyrBiasN = 0.00003f;
    float imuGravity = 9.80511f, imuRPYWeight = 0.01f;
    bool correct = false;
    std::vector<double> extRotV, extRPYV, extTransV;
    // Mapping/*
    float z_tollerance = FLT_MAX, rotation_tollerance = FLT_MAX;
    std::string regMethod = "DIRECT1";
    float ndtResolution = 1.0f, ndtEpsilon = 0.01f;
    float timeInterval = 0.2f;
    double mappingProcessInterval = 0.15;
    float mappingSurfLeafSize = 0.2f, surroundingKeyframeMapLeafSize = 0.4f;
    float surroundingkeyframeAddingDistThreshold = 1.0f, surroundingkeyframeAddingAngleThreshold = 0.2f;
    float surroundingKeyframeDensity = 1.0f, surroundingKeyframeSearchRadius = 50.0f;
    float globalMapVisualizationSearchRadius = 1e3f, globalMapVisualizationPoseDensity = 10.0f, globalMapVisualizationLeafSize = 1.0f;

    std::string param_error;     // non-empty when the yaml could not be used (the reference calls ros::shutdown())

    ParamServer() = default;
    explicit ParamServer(const std::string& yaml_path) { load(yaml_path); }

    bool load(const std::string& yaml_path) {
        std::map<std::string, std::string> kv;
        if (!parse_yaml(yaml_path, kv)) { param_error = "cannot read " + yaml_path; return false; }
        auto S = [&](const char* k, std::string& v) { auto it = kv.find(k); if (it != kv.end()) v = it->second; };
        auto F = [&](const char* k, float& v) { auto it = kv.find(k); if (it != kv.end()) v = (float)std::atof(it->second.c_str()); };
        auto D = [&](const char* k, double& v) { auto it = kv.find(k); if (it != kv.end()) v = std::atof(it->second.c_str()); };
        auto I = [&](const char* k, int& v) { auto it = kv.find(k); if (it != kv.end()) v = std::atoi(it->second.c_str()); };
        auto B = [&](const char* k, bool& v) { auto it = kv.find(k); if (it != kv.end()) v = (it->second == "true" || it->second == "True" || it->second == "1"); };
        auto L = [&](const char* k, std::vector<double>& v) {
            auto it = kv.find(k); if (it == kv.end()) return;
            v.clear(); std::string s = it->second;
            for (char& c : s) if (c == '[' || c == ']' || c == ',') c = ' ';
            std::stringstream ss(s); double d; while (ss >> d) v.push_back(d);
        };
        S("System/pointCloudTopic", pointCloudTopic); S("System/imuTopic", imuTopic); S("System/odomTopic", odomTopic);
        std::string modeStr = "lio"; S("System/mode", modeStr);
        if (modeStr == "lio") mode = ModeType::LIO;
        else if (modeStr == "relo") mode = ModeType::RELO;
        else { param_error = "Invalid Mode Type (must be either 'lio' or 'relo'): " + modeStr; return false; }
        I("System/numberOfCores", numberOfCores);
        S("System/lidarFrame", lidarFrame); S("System/baselinkFrame", baselinkFrame);
        S("System/odometryFrame", odometryFrame); S("System/mapFrame", mapFrame);
        S("System/savePCDDirectory", savePCDDirectory); S("System/saveSessionDirectory", saveSessionDirectory);
        std::string sensorStr = " "; S("Sensors/sensor", sensorStr);
        if (sensorStr == "velodyne") sensor = SensorType::VELODYNE;
        else if (sensorStr == "ouster") sensor = SensorType::OUSTER;
        else if (sensorStr == "livox") sensor = SensorType::LIVOX;
        else if (sensorStr == "robosense") sensor = SensorType::ROBOSENSE;
        else if (sensorStr == "mulran") sensor = SensorType::MULRAN;
        else { param_error = "Invalid Sensor Type: " + sensorStr; return false; }
        I("Sensors/N_SCAN", N_SCAN); I("Sensors/Horizon_SCAN", Horizon_SCAN);
        B("Sensors/have_ring_time_channel", have_ring_time_channel);
        I("Sensors/downsampleRate", downsampleRate); I("Sensors/point_filter_num", point_filter_num);
        F("Sensors/lidarMinRange", lidarMinRange); F("Sensors/lidarMaxRange", lidarMaxRange);
        I("Sensors/imuType", imuType); F("Sensors/imuRate", imuRate);
        F("Sensors/imuAccNoise", imuAccNoise); F("Sensors/imuGyrNoise", imuGyrNoise);
        F("Sensors/imuAccBiasN", imuAccBiasN); F("Sensors/imuGyrBiasN", imuGyrBiasN);
        F("Sensors/imuGravity", imuGravity); F("Sensors/imuRPYWeight", imuRPYWeight);
        B("Sensors/correct", correct);
        L("Sensors/extrinsicRot", extRotV); L("Sensors/extrinsicRPY", extRPYV); L("Sensors/extrinsicTrans", extTransV);
        F("Mapping/z_tollerance", z_tollerance); F("Mapping/rotation_tollerance", rotation_tollerance);
        S("Mapping/regMethod", regMethod); F("Mapping/ndtResolution", ndtResolution); F("Mapping/ndtEpsilon", ndtEpsilon);
        F("Mapping/timeInterval", timeInterval); D("Mapping/mappingProcessInterval", mappingProcessInterval);
        F("Mapping/mappingSurfLeafSize", mappingSurfLeafSize); F("Mapping/surroundingKeyframeMapLeafSize", surroundingKeyframeMapLeafSize);
        F("Mapping/surroundingkeyframeAddingDistThreshold", surroundingkeyframeAddingDistThreshold);
        F("Mapping/surroundingkeyframeAddingAngleThreshold", surroundingkeyframeAddingAngleThreshold);
        F("Mapping/surroundingKeyframeDensity", surroundingKeyframeDensity);
        F("Mapping/surroundingKeyframeSearchRadius", surroundingKeyframeSearchRadius);
        F("Mapping/globalMapVisualizationSearchRadius", globalMapVisualizationSearchRadius);
        F("Mapping/globalMapVisualizationPoseDensity", globalMapVisualizationPoseDensity);
        F("Mapping/globalMapVisualizationLeafSize", globalMapVisualizationLeafSize);
        return true;
    }

    // flat "Section/key" -> raw value string
    static bool parse_yaml(const std::string& path, std::map<std::string, std::string>& out) {
        std::ifstream f(path);
        if (!f) return false;
        std::string line, section, pending_key, pending_val;
        int depth = 0;
        auto strip = [](std::string s) {
            size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
            return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
        };
        auto unquote = [](std::string s) {
            if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
            return s;
        };
        while (std::getline(f, line)) {
            bool in_q = false; size_t cut = std::string::npos;            // strip comments outside quotes
            for (size_t i = 0; i < line.size(); ++i) { if (line[i] == '"') in_q = !in_q; if (line[i] == '#' && !in_q) { cut = i; break; } }
            if (cut != std::string::npos) line = line.substr(0, cut);
            if (strip(line).empty()) continue;
            if (depth > 0) {                                              // continuation of a multi-line inline list
                pending_val += " " + strip(line);
                for (char c : line) { if (c == '[') ++depth; if (c == ']') --depth; }
                if (depth == 0) out[pending_key] = pending_val;
                continue;
            }
            const size_t indent = line.find_first_not_of(" \t");
            const size_t colon = line.find(':');
            if (colon == std::string::npos) continue;
            const std::string key = strip(line.substr(0, colon));
            const std::string val = strip(line.substr(colon + 1));
            if (indent == 0) { section = key; if (!val.empty()) out[key] = unquote(val); continue; }
            const std::string full = section + "/" + key;
            for (char c : val) { if (c == '[') ++depth; if (c == ']') --depth; }
            if (depth > 0) { pending_key = full; pending_val = val; continue; }
            out[full] = unquote(val);
        }
        return true;
    }
};

}  // namespace liloc_host
