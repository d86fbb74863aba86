// crv/grid3d.h -- dense 3D grid + the reference's ".xyz" cloud file format.
//
// Same public surface as Grid3D<T> (Grid3D.h:5-192): Load / operator() / GetByteSize /
// GetElementSize / GetSize / GetVoxelSize / GetVoxelCount / GetData / Copy / Save / GetMajorant.
// Storage is x-fastest: idx = x + y*nx + z*nx*ny (Grid3D.h:112-116), which is also what gets
// uploaded tightly packed into the 3D image (main.cpp:288-293).
//
// File format (GridData.h:1-25): 3 x u32 counts, 3 x f64 voxel sizes, then floats in ZYX order
// (z fastest, x slowest).  Load transposes to x-fastest (Grid3D.h:73-85).  Save writes the raw
// x-fastest array (Grid3D.h:167-178) -- i.e. Save is NOT the inverse of Load in the reference; that
// asymmetry is kept, and SaveZFastest() is provided as the writer Load actually expects.
#pragma once

#include <array>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace crv {

template <typename T>
class Grid3D {
public:
    // Grid3D.h:45-90; returns nullptr when the file cannot be opened (caller owns the result)
    static Grid3D<T>* Load(const std::string& filename)
    {
        std::ifstream in(filename, std::ifstream::in | std::ifstream::binary);
        if (!in.good()) return nullptr;
        uint32_t n[3] = {0, 0, 0};
        double vs[3] = {0, 0, 0};
        in.read(reinterpret_cast<char*>(n), sizeof(n));
        in.read(reinterpret_cast<char*>(vs), sizeof(vs));
        Grid3D<T>* grid = new Grid3D<T>(n[0], n[1], n[2], vs[0], vs[1], vs[2]);
        // the file is z-fastest; read one z-run at a time and scatter it with stride nx*ny
        std::vector<float> run(n[2]);
        const size_t plane = static_cast<size_t>(n[0]) * n[1];
        for (uint32_t x = 0; x < n[0]; ++x)
            for (uint32_t y = 0; y < n[1]; ++y) {
                in.read(reinterpret_cast<char*>(run.data()), sizeof(float) * n[2]);
                T* dst = grid->m_data.data() + x + static_cast<size_t>(n[0]) * y;
                for (uint32_t z = 0; z < n[2]; ++z) dst[plane * z] = static_cast<T>(run[z]);
            }
        return grid;
    }

    Grid3D(unsigned sizeX, unsigned sizeY, unsigned sizeZ, double voxelSizeX = 1, double voxelSizeY = 1, double voxelSizeZ = 1)
        : m_count{sizeX, sizeY, sizeZ}, m_voxel{voxelSizeX, voxelSizeY, voxelSizeZ}
    {
        m_data.resize(static_cast<size_t>(sizeX) * sizeY * sizeZ);
    }

    // Grid3D.h:112-121: bounds-checked, throws std::out_of_range
    T& operator()(size_t x, size_t y, size_t z)
    {
        if (x < m_count[0] && y < m_count[1] && z < m_count[2])
            return m_data[x + y * m_count[0] + z * static_cast<size_t>(m_count[0]) * m_count[1]];
        throw std::out_of_range("Index out of 3D grid bounds");
    }

    size_t GetByteSize() const { return m_data.size() * sizeof(T); }
    size_t GetElementSize() const { return sizeof(T); }
    size_t GetSize() const { return m_data.size(); }
    std::array<double, 3> GetVoxelSize() const { return m_voxel; }
    std::array<uint32_t, 3> GetVoxelCount() const { return m_count; }
    void* GetData() { return m_data.data(); }
    const T* data() const { return m_data.data(); }

    // Grid3D.h:159-164 (resizes to `size` ELEMENTS and copies `size` BYTES, as the reference does)
    void Copy(void* src, size_t size)
    {
        m_data.resize(size);
        std::memcpy(m_data.data(), src, size);
    }

    // Grid3D.h:167-178: header + raw x-fastest data
    void Save(const std::string& filename) const
    {
        std::ofstream out(filename, std::ofstream::out | std::ofstream::binary);
        WriteHeader(out);
        out.write(reinterpret_cast<const char*>(m_data.data()), sizeof(T) * m_data.size());
    }

    // GridData::save convention (GridData.h:12-25): z-fastest on disk -- round-trips through Load
    void SaveZFastest(const std::string& filename) const
    {
        std::ofstream out(filename, std::ofstream::out | std::ofstream::binary);
        WriteHeader(out);
        std::vector<float> run(m_count[2]);
        const size_t plane = static_cast<size_t>(m_count[0]) * m_count[1];
        for (uint32_t x = 0; x < m_count[0]; ++x)
            for (uint32_t y = 0; y < m_count[1]; ++y) {
                const T* src = m_data.data() + x + static_cast<size_t>(m_count[0]) * y;
                for (uint32_t z = 0; z < m_count[2]; ++z) run[z] = static_cast<float>(src[plane * z]);
                out.write(reinterpret_cast<const char*>(run.data()), sizeof(float) * m_count[2]);
            }
    }

    // Grid3D.h:181-192: max over all voxels, starting from T() (so never below zero)
    T GetMajorant() const
    {
        T majorant = T();
        for (const T& v : m_data)
            if (majorant < v) majorant = v;
        return majorant;
    }

private:
    void WriteHeader(std::ofstream& out) const
    {
        out.write(reinterpret_cast<const char*>(m_count.data()), sizeof(uint32_t) * 3);
        out.write(reinterpret_cast<const char*>(m_voxel.data()), sizeof(double) * 3);
    }

    std::vector<T> m_data;
    std::array<uint32_t, 3> m_count;
    std::array<double, 3> m_voxel;
};

}  // namespace crv
