// liloc_b200/host/session_io.cpp -- see session_io.hpp.
#include "session_io.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <map>
#include <sstream>

namespace liloc_host {

namespace {

struct PcdField { std::string name; int size = 4; char type = 'F'; int count = 1; int offset = 0; };

std::vector<std::string> split_ws(const std::string& s) {
    std::vector<std::string> out;
    std::istringstream is(s);
    std::string t;
    while (is >> t) out.push_back(t);
    return out;
}

// value of one scalar of a binary record as float
float read_scalar(const unsigned char* p, const PcdField& f) {
    switch (f.type) {
        case 'F': if (f.size == 4) { float v; std::memcpy(&v, p, 4); return v; } else { double v; std::memcpy(&v, p, 8); return (float)v; }
        case 'U': if (f.size == 1) return (float)*p; if (f.size == 2) { unsigned short v; std::memcpy(&v, p, 2); return (float)v; } { unsigned v; std::memcpy(&v, p, 4); return (float)v; }
        default:  if (f.size == 1) return (float)*(const signed char*)p; if (f.size == 2) { short v; std::memcpy(&v, p, 2); return (float)v; } { int v; std::memcpy(&v, p, 4); return (float)v; }
    }
}

// liblzf's lzf_decompress (the codec behind PCL's binary_compressed): control byte < 32 = literal run of ctrl + 1
// bytes; otherwise a back reference of length (ctrl >> 5) + 2 (+ one extension byte when the 3-bit field is 7) at
// distance ((ctrl & 31) << 8 | next byte) + 1.  Returns false on a malformed stream or a size mismatch.
bool lzf_decompress(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len) {
    size_t ip = 0, op = 0;
    while (ip < in_len) {
        unsigned ctrl = in[ip++];
        if (ctrl < 32) {
            ++ctrl;
            if (ip + ctrl > in_len || op + ctrl > out_len) return false;
            std::memcpy(out + op, in + ip, ctrl);
            ip += ctrl; op += ctrl;
        } else {
            size_t len = ctrl >> 5;
            if (len == 7) { if (ip >= in_len) return false; len += in[ip++]; }
            if (ip >= in_len) return false;
            const size_t dist = ((size_t)(ctrl & 31u) << 8 | in[ip++]) + 1;
            len += 2;
            if (dist > op || op + len > out_len) return false;
            for (size_t k = 0; k < len; ++k, ++op) out[op] = out[op - dist];     // may overlap: byte by byte
        }
    }
    return op == out_len;
}

// gtsam::Rot3(Quaternion(w, x, y, z)).roll() / pitch() / yaw()  (R = Rz(yaw) Ry(pitch) Rx(roll))
void quat_to_rpy(double qx, double qy, double qz, double qw, float& roll, float& pitch, float& yaw) {
    const double n = std::sqrt(qx * qx + qy * qy + qz * qz + qw * qw);
    if (n > 0) { qx /= n; qy /= n; qz /= n; qw /= n; }
    const double r00 = 1 - 2 * (qy * qy + qz * qz), r10 = 2 * (qx * qy + qw * qz);
    const double r20 = 2 * (qx * qz - qw * qy), r21 = 2 * (qy * qz + qw * qx), r22 = 1 - 2 * (qx * qx + qy * qy);
    roll = (float)std::atan2(r21, r22);
    pitch = (float)std::atan2(-r20, std::sqrt(r21 * r21 + r22 * r22));
    yaw = (float)std::atan2(r10, r00);
}

// gtsam::Rot3::RzRyRx(roll, pitch, yaw).toQuaternion()
void rpy_to_quat(double roll, double pitch, double yaw, double q[4] /*x y z w*/) {
    const double cr = std::cos(roll * 0.5), sr = std::sin(roll * 0.5), cp = std::cos(pitch * 0.5), sp = std::sin(pitch * 0.5);
    const double cy = std::cos(yaw * 0.5), sy = std::sin(yaw * 0.5);
    q[0] = sr * cp * cy - cr * sp * sy; q[1] = cr * sp * cy + sr * cp * sy; q[2] = cr * cp * sy - sr * sp * cy; q[3] = cr * cp * cy + sr * sp * sy;
}

}  // namespace

bool loadPCD(const std::string& path, std::vector<liloc_point>& out, std::string& err) {
    out.clear();
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open " + path; return false; }
    std::vector<PcdField> fields;
    long long points = -1, width = -1, height = 1;
    std::string data;
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == '#') continue;
        const std::vector<std::string> t = split_ws(line);
        if (t.empty()) continue;
        if (t[0] == "FIELDS") { fields.resize(t.size() - 1); for (size_t i = 1; i < t.size(); ++i) fields[i - 1].name = t[i]; }
        else if (t[0] == "SIZE")  { for (size_t i = 1; i < t.size() && i - 1 < fields.size(); ++i) fields[i - 1].size = std::atoi(t[i].c_str()); }
        else if (t[0] == "TYPE")  { for (size_t i = 1; i < t.size() && i - 1 < fields.size(); ++i) fields[i - 1].type = t[i][0]; }
        else if (t[0] == "COUNT") { for (size_t i = 1; i < t.size() && i - 1 < fields.size(); ++i) fields[i - 1].count = std::atoi(t[i].c_str()); }
        else if (t[0] == "WIDTH" && t.size() > 1) width = std::atoll(t[1].c_str());
        else if (t[0] == "HEIGHT" && t.size() > 1) height = std::atoll(t[1].c_str());
        else if (t[0] == "POINTS" && t.size() > 1) points = std::atoll(t[1].c_str());
        else if (t[0] == "DATA" && t.size() > 1) { data = t[1]; break; }
    }
    if (fields.empty() || data.empty()) { err = path + ": not a PCD file (no FIELDS / DATA header)"; return false; }
    if (points < 0) points = width * height;
    if (points < 0 || points > 0x7fffffffLL) { err = path + ": bad POINTS"; return false; }
    int rec = 0, ix = -1, iy = -1, iz = -1, ii = -1;
    for (size_t k = 0; k < fields.size(); ++k) {
        PcdField& fd = fields[k];
        if (fd.size <= 0 || fd.count <= 0) { err = path + ": bad SIZE / COUNT"; return false; }
        fd.offset = rec; rec += fd.size * fd.count;
        if (fd.name == "x") ix = (int)k; else if (fd.name == "y") iy = (int)k; else if (fd.name == "z") iz = (int)k;
        else if (fd.name == "intensity") ii = (int)k;
    }
    if (ix < 0 || iy < 0 || iz < 0) { err = path + ": no x / y / z fields"; return false; }
    out.resize((size_t)points);
    if (data == "binary") {
        std::vector<unsigned char> buf((size_t)points * rec);
        f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)buf.size());
        if ((long long)f.gcount() != (long long)buf.size()) { err = path + ": truncated binary data"; return false; }
        for (long long p = 0; p < points; ++p) {
            const unsigned char* r = buf.data() + (size_t)p * rec;
            out[p].x = read_scalar(r + fields[ix].offset, fields[ix]);
            out[p].y = read_scalar(r + fields[iy].offset, fields[iy]);
            out[p].z = read_scalar(r + fields[iz].offset, fields[iz]);
            out[p].intensity = ii >= 0 ? read_scalar(r + fields[ii].offset, fields[ii]) : 0.f;
        }
    } else if (data == "ascii") {
        for (long long p = 0; p < points; ++p) {
            if (!std::getline(f, line)) { err = path + ": truncated ascii data"; return false; }
            const std::vector<std::string> t = split_ws(line);
            size_t col = 0;
            float v[4] = {0, 0, 0, 0};
            for (size_t k = 0; k < fields.size(); ++k) {
                for (int c = 0; c < fields[k].count; ++c, ++col) {
                    if (col >= t.size()) { err = path + ": short ascii row"; return false; }
                    if (c == 0) {
                        const float val = (float)std::atof(t[col].c_str());
                        if ((int)k == ix) v[0] = val; else if ((int)k == iy) v[1] = val; else if ((int)k == iz) v[2] = val; else if ((int)k == ii) v[3] = val;
                    }
                }
            }
            out[p] = liloc_point{v[0], v[1], v[2], v[3]};
        }
    } else if (data == "binary_compressed") {
        // pcl::PCDWriter::writeBinaryCompressed: uint32 compressed size, uint32 uncompressed size, LZF stream; the
        // uncompressed block is field-major (all x, then all y, ...), each field SIZE * COUNT bytes per point.
        unsigned sizes[2];
        f.read(reinterpret_cast<char*>(sizes), 8);
        if (f.gcount() != 8) { err = path + ": truncated binary_compressed header"; return false; }
        const size_t csize = sizes[0], usize = sizes[1];
        if (usize != (size_t)points * rec) { err = path + ": binary_compressed size does not match POINTS x record size"; return false; }
        std::vector<unsigned char> cbuf(csize), buf(usize);
        f.read(reinterpret_cast<char*>(cbuf.data()), (std::streamsize)csize);
        if ((size_t)f.gcount() != csize) { err = path + ": truncated binary_compressed data"; return false; }
        if (!lzf_decompress(cbuf.data(), csize, buf.data(), usize)) { err = path + ": corrupt LZF stream"; return false; }
        std::vector<size_t> col(fields.size());
        size_t acc = 0;
        for (size_t k = 0; k < fields.size(); ++k) { col[k] = acc; acc += (size_t)fields[k].size * fields[k].count * (size_t)points; }
        auto at = [&](int k, long long p) { return read_scalar(buf.data() + col[k] + (size_t)p * fields[k].size * fields[k].count, fields[k]); };
        for (long long p = 0; p < points; ++p) {
            out[p].x = at(ix, p); out[p].y = at(iy, p); out[p].z = at(iz, p);
            out[p].intensity = ii >= 0 ? at(ii, p) : 0.f;
        }
    } else {
        err = path + ": DATA " + data + " is not supported (ascii, binary and binary_compressed are; the reference writes binary)";
        return false;
    }
    return true;
}

bool savePCDBinary(const std::string& path, const liloc_point* pts, int n, std::string& err) {
    std::ofstream f(path, std::ios::binary);
    if (!f) { err = "cannot write " + path; return false; }
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
      << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
    if (n > 0) f.write(reinterpret_cast<const char*>(pts), (std::streamsize)n * (std::streamsize)sizeof(liloc_point));
    if (!f) { err = "write failed: " + path; return false; }
    return true;
}

bool savePosesPCDBinary(const std::string& path, const std::vector<PointTypePose>& kp, std::string& err) {
    std::ofstream f(path, std::ios::binary);
    if (!f) { err = "cannot write " + path; return false; }
    const size_t n = kp.size();
    f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z _ intensity roll pitch yaw time _\n"
      << "SIZE 4 4 4 1 4 4 4 4 8 1\nTYPE F F F U F F F F F U\nCOUNT 1 1 1 4 1 1 1 1 1 8\n"
      << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
    for (const PointTypePose& p : kp) {
        unsigned char rec[48] = {0};
        const float head[3] = {p.x, p.y, p.z};
        const float one = 1.0f;                               // PCL_ADD_POINT4D: data[3] = 1
        const float mid[4] = {p.intensity, p.roll, p.pitch, p.yaw};
        std::memcpy(rec, head, 12); std::memcpy(rec + 12, &one, 4); std::memcpy(rec + 16, mid, 16); std::memcpy(rec + 32, &p.time, 8);
        f.write(reinterpret_cast<const char*>(rec), 48);
    }
    if (!f) { err = "write failed: " + path; return false; }
    return true;
}

bool loadSessionGraph(const std::string& g2o_path, std::vector<PointTypePose>& keyPoses, std::string& err) {
    keyPoses.clear();
    std::ifstream f(g2o_path);
    if (!f) { err = "cannot open " + g2o_path; return false; }
    std::multimap<int, PointTypePose> nodes;          // SessionNodes: ascending id, duplicates kept (dataLoader.hpp:47)
    std::string line;
    while (std::getline(f, line)) {
        const std::vector<std::string> t = split_ws(line);
        if (t.size() < 9 || t[0] != "VERTEX_SE3:QUAT") continue;
        const int id = std::atoi(t[1].c_str());
        PointTypePose p{};
        p.x = (float)std::atof(t[2].c_str()); p.y = (float)std::atof(t[3].c_str()); p.z = (float)std::atof(t[4].c_str());
        quat_to_rpy(std::atof(t[5].c_str()), std::atof(t[6].c_str()), std::atof(t[7].c_str()), std::atof(t[8].c_str()), p.roll, p.pitch, p.yaw);
        p.intensity = (float)id; p.time = 0.0;
        nodes.insert({id, p});
    }
    for (const auto& kv : nodes) keyPoses.push_back(kv.second);
    if (keyPoses.empty()) { err = g2o_path + ": no VERTEX_SE3:QUAT lines"; return false; }
    return true;
}

bool saveSessionGraph(const std::string& g2o_path, const std::vector<PointTypePose>& kp, std::string& err) {
    std::ofstream f(g2o_path);
    if (!f) { err = "cannot write " + g2o_path; return false; }
    std::vector<std::string> vertices, edges;
    for (size_t i = 0; i < kp.size(); ++i) {
        double q[4];
        rpy_to_quat(kp[i].roll, kp[i].pitch, kp[i].yaw, q);
        vertices.push_back("VERTEX_SE3:QUAT " + std::to_string(i) + " " + std::to_string((double)kp[i].x) + " " + std::to_string((double)kp[i].y) + " " +
                           std::to_string((double)kp[i].z) + " " + std::to_string(q[0]) + " " + std::to_string(q[1]) + " " + std::to_string(q[2]) + " " + std::to_string(q[3]));
        if (i > 0) {                                  // poseFrom.between(poseTo) (:574-581)
            const Affine3f a = getTransformation(kp[i - 1].x, kp[i - 1].y, kp[i - 1].z, kp[i - 1].roll, kp[i - 1].pitch, kp[i - 1].yaw);
            const Affine3f b = getTransformation(kp[i].x, kp[i].y, kp[i].z, kp[i].roll, kp[i].pitch, kp[i].yaw);
            const Affine3f r = a.inverse() * b;
            float x, y, z, ro, pi, ya;
            getTranslationAndEulerAngles(r, x, y, z, ro, pi, ya);
            rpy_to_quat(ro, pi, ya, q);
            edges.push_back("EDGE_SE3:QUAT " + std::to_string(i - 1) + " " + std::to_string(i) + " " + std::to_string((double)x) + " " + std::to_string((double)y) + " " +
                            std::to_string((double)z) + " " + std::to_string(q[0]) + " " + std::to_string(q[1]) + " " + std::to_string(q[2]) + " " + std::to_string(q[3]));
        }
    }
    for (const auto& l : vertices) f << l << "\n";
    for (const auto& l : edges) f << l << "\n";
    if (!f) { err = "write failed: " + g2o_path; return false; }
    return true;
}

bool loadKeyClouds(const std::string& pcd_dir, std::vector<std::vector<liloc_point>>& clouds, std::string& err) {
    clouds.clear();
    namespace fs = std::filesystem;
    std::error_code ec;
    if (!fs::is_directory(pcd_dir, ec)) { err = "not a directory: " + pcd_dir; return false; }
    std::vector<std::pair<int, std::string>> names;
    for (const auto& e : fs::directory_iterator(pcd_dir, ec)) {
        const std::string stem = e.path().stem().string();
        if (e.path().extension() != ".pcd" || stem.empty() || !std::all_of(stem.begin(), stem.end(), [](char c) { return c >= '0' && c <= '9'; })) continue;
        names.push_back({std::atoi(stem.c_str()), e.path().string()});          // fileNameSort: std::stoi of the stem
    }
    std::sort(names.begin(), names.end());
    clouds.resize(names.size());
    for (size_t i = 0; i < names.size(); ++i) if (!loadPCD(names[i].second, clouds[i], err)) return false;
    return true;
}

}  // namespace liloc_host
