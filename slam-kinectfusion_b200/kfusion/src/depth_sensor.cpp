// depth_sensor.cpp -- dataset frame source (reference: kfusion/src/depth_sensor.cpp:11-46 open, :186-196 getFrame)
// and the PNG decoder it needs (the reference calls cv::imread; OpenCV's C++ library is not part of this build).
// Host-only code: nothing here touches the GPU.
#include "depth_sensor.h"
#include "kinectfusion.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <dirent.h>
#include <zlib.h>

using namespace cv; // CV_8UC3 / CV_32FC1 are enumerators of cvlite's namespace, macros with real OpenCV

namespace
{
typedef unsigned char u8;
typedef unsigned int u32;

u32 be32(const u8 *p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | (u32)p[3]; }
void put32(std::vector<u8> &v, u32 x)
{
    v.push_back((u8)(x >> 24)); v.push_back((u8)(x >> 16)); v.push_back((u8)(x >> 8)); v.push_back((u8)x);
}
bool fail(std::string *err, const std::string &why)
{
    if (err) *err = why;
    return false;
}
bool slurp(const std::string &path, std::vector<u8> &out)
{
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const size_t got = out.empty() ? 0 : std::fread(out.data(), 1, out.size(), f);
    std::fclose(f);
    return got == out.size();
}
int paeth(int a, int b, int c)
{
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
void chunk(std::vector<u8> &file, const char type[4], const std::vector<u8> &data)
{
    put32(file, (u32)data.size());
    const size_t at = file.size();
    file.insert(file.end(), type, type + 4);
    file.insert(file.end(), data.begin(), data.end());
    put32(file, (u32)crc32(0L, file.data() + at, (uInt)(file.size() - at)));
}
bool write_png(const std::string &path, const std::vector<u8> &rows, int width, int height, int bit_depth, int color_type)
{
    static const u8 sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<u8> file(sig, sig + 8), ihdr, idat;
    put32(ihdr, (u32)width); put32(ihdr, (u32)height);
    ihdr.push_back((u8)bit_depth); ihdr.push_back((u8)color_type); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(file, "IHDR", ihdr);
    uLongf n = compressBound((uLong)rows.size());
    idat.resize(n);
    if (compress2(idat.data(), &n, rows.data(), (uLong)rows.size(), 6) != Z_OK) return false;
    idat.resize(n);
    chunk(file, "IDAT", idat);
    chunk(file, "IEND", std::vector<u8>());
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = std::fwrite(file.data(), 1, file.size(), f) == file.size();
    std::fclose(f);
    return ok;
}
} // namespace

bool kf::png::read(const std::string &path, Image &out, std::string *err)
{
    std::vector<u8> file;
    if (!slurp(path, file)) return fail(err, "cannot read " + path);
    static const u8 sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0) return fail(err, path + ": not a PNG file");

    int width = 0, height = 0, bit_depth = 0, color_type = -1;
    std::vector<u8> zdata, palette;
    bool seen_end = false, interlaced = false;
    for (size_t at = 8; at + 12 <= file.size() && !seen_end;)
    {
        const u32 len = be32(&file[at]);
        if ((size_t)len > file.size() - at - 12) return fail(err, path + ": truncated chunk");
        const u8 *type = &file[at + 4], *data = &file[at + 8];
        if (be32(data + len) != (u32)crc32(0L, type, len + 4)) return fail(err, path + ": chunk CRC mismatch");
        if (!std::memcmp(type, "IHDR", 4))
        {
            if (len != 13) return fail(err, path + ": bad IHDR");
            width = (int)be32(data); height = (int)be32(data + 4);
            bit_depth = data[8]; color_type = data[9];
            if (data[10] != 0 || data[11] != 0) return fail(err, path + ": unknown compression / filter method");
            if (data[12] > 1) return fail(err, path + ": unknown interlace method");
            interlaced = data[12] == 1;
        }
        else if (!std::memcmp(type, "PLTE", 4)) palette.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) zdata.insert(zdata.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) seen_end = true;
        at += 12 + (size_t)len;
    }
    if (color_type < 0 || !seen_end) return fail(err, path + ": missing IHDR or IEND");
    if (width <= 0 || height <= 0 || width > (1 << 16) || height > (1 << 16)) return fail(err, path + ": unreasonable size");
    int file_channels;
    switch (color_type)
    {
    case 0: file_channels = 1; break;
    case 2: file_channels = 3; break;
    case 3: file_channels = 1; break;
    case 4: file_channels = 2; break;
    case 6: file_channels = 4; break;
    default: return fail(err, path + ": bad colour type");
    }
    const bool depth_ok = (color_type == 0 && (bit_depth == 1 || bit_depth == 2 || bit_depth == 4 || bit_depth == 8 || bit_depth == 16)) ||
                          (color_type == 3 && (bit_depth == 1 || bit_depth == 2 || bit_depth == 4 || bit_depth == 8)) ||
                          ((color_type == 2 || color_type == 4 || color_type == 6) && (bit_depth == 8 || bit_depth == 16));
    if (!depth_ok) return fail(err, path + ": bad bit depth for its colour type");
    if (color_type == 3 && (palette.empty() || palette.size() % 3)) return fail(err, path + ": palette image without a valid PLTE");

    // Sub-images in file order: the whole image, or the seven Adam7 passes (PNG spec 8.2): pixel i of row j of a
    // pass lands at (x0 + i dx, y0 + j dy).  Each is filtered on its own: rows of (filter byte + stride bytes).
    struct Pass { int x0, y0, dx, dy; };
    static const Pass whole[1] = {{0, 0, 1, 1}};
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    const Pass *passes = interlaced ? adam7 : whole;
    const int npasses = interlaced ? 7 : 1;
    const size_t bits_px = (size_t)file_channels * bit_depth;
    size_t total = 0;
    for (int k = 0; k < npasses; ++k)
    {
        const int pw = (width - passes[k].x0 + passes[k].dx - 1) / passes[k].dx, ph = (height - passes[k].y0 + passes[k].dy - 1) / passes[k].dy;
        if (pw > 0 && ph > 0) total += (((size_t)pw * bits_px + 7) / 8 + 1) * (size_t)ph;
    }
    std::vector<u8> raw(total);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, zdata.data(), (uLong)zdata.size()) != Z_OK || raw_len != raw.size())
        return fail(err, path + ": corrupt image data");

    out = Image();
    out.width = width; out.height = height;
    out.channels = color_type == 3 ? 3 : file_channels;
    out.bit_depth = bit_depth == 16 ? 16 : 8;
    const size_t n = (size_t)width * height * out.channels;
    if (bit_depth == 16) out.data16.resize(n);
    else out.data8.resize(n);

    const size_t bpp = std::max<size_t>(1, bits_px / 8); // bytes per complete pixel, at least 1 (PNG spec 9.2)
    size_t at = 0;
    for (int k = 0; k < npasses; ++k)
    {
        const Pass &ps = passes[k];
        const int pw = (width - ps.x0 + ps.dx - 1) / ps.dx, ph = (height - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw <= 0 || ph <= 0) continue;
        const size_t stride = ((size_t)pw * bits_px + 7) / 8;
        for (int j = 0; j < ph; ++j)
        {
            // undo the row filter in place
            u8 *row = &raw[at + (stride + 1) * (size_t)j + 1];
            const u8 *up = j ? row - (stride + 1) : nullptr;
            const int filter = row[-1];
            if (filter > 4) return fail(err, path + ": bad row filter");
            for (size_t i = 0; i < stride; ++i)
            {
                const int a = i >= bpp ? row[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= bpp) ? up[i - bpp] : 0;
                int pred = 0;
                if (filter == 1) pred = a;
                else if (filter == 2) pred = b;
                else if (filter == 3) pred = (a + b) >> 1;
                else if (filter == 4) pred = paeth(a, b, c);
                row[i] = (u8)(row[i] + pred);
            }
            // scatter the row's pixels
            const int y = ps.y0 + j * ps.dy;
            for (int i = 0; i < pw; ++i)
            {
                const size_t o = ((size_t)y * width + (ps.x0 + i * ps.dx)) * out.channels;
                if (bit_depth == 16)
                    for (int ch = 0; ch < file_channels; ++ch)
                        out.data16[o + ch] = (unsigned short)((row[2 * ((size_t)i * file_channels + ch)] << 8) | row[2 * ((size_t)i * file_channels + ch) + 1]);
                else if (bit_depth == 8 && color_type != 3)
                    for (int ch = 0; ch < file_channels; ++ch) out.data8[o + ch] = row[(size_t)i * file_channels + ch];
                else
                {
                    // 1/2/4/8-bit sample i of the row, most significant bits first
                    const int per = 8 / bit_depth, shift = (per - 1 - i % per) * bit_depth;
                    const int v = (row[i / per] >> shift) & ((1 << bit_depth) - 1);
                    if (color_type == 3)
                    {
                        if ((size_t)v * 3 + 2 >= palette.size()) return fail(err, path + ": palette index out of range");
                        out.data8[o] = palette[3 * v]; out.data8[o + 1] = palette[3 * v + 1]; out.data8[o + 2] = palette[3 * v + 2];
                    }
                    else out.data8[o] = (u8)(v * 255 / ((1 << bit_depth) - 1)); // grey 1/2/4 -> 8 bits, libpng's expansion
                }
            }
        }
        at += (stride + 1) * (size_t)ph;
    }
    return true;
}

bool kf::png::write_gray16(const std::string &path, const unsigned short *pix, int width, int height)
{
    std::vector<u8> rows;
    rows.reserve(((size_t)width * 2 + 1) * height);
    for (int y = 0; y < height; ++y)
    {
        rows.push_back(0); // filter: none
        for (int x = 0; x < width; ++x)
        {
            const unsigned short v = pix[(size_t)y * width + x];
            rows.push_back((u8)(v >> 8)); rows.push_back((u8)v);
        }
    }
    return write_png(path, rows, width, height, 16, 0);
}
bool kf::png::write_rgb8(const std::string &path, const unsigned char *rgb, int width, int height)
{
    std::vector<u8> rows;
    rows.reserve(((size_t)width * 3 + 1) * height);
    for (int y = 0; y < height; ++y)
    {
        rows.push_back(0);
        rows.insert(rows.end(), rgb + (size_t)y * width * 3, rgb + (size_t)(y + 1) * width * 3);
    }
    return write_png(path, rows, width, height, 8, 2);
}

// ---- the frame source -----------------------------------------------------------------------------------
namespace
{
// cv::glob(dir + "/*.png"): the matching names, sorted
std::vector<std::string> list_png(const std::string &dir)
{
    std::vector<std::string> names;
    if (DIR *d = opendir(dir.c_str()))
    {
        while (const dirent *e = readdir(d))
        {
            const std::string n = e->d_name;
            if (n.size() > 4 && n.compare(n.size() - 4, 4, ".png") == 0) names.push_back(dir + "/" + n);
        }
        closedir(d);
    }
    std::sort(names.begin(), names.end());
    return names;
}
// cv::imread(name, 1): always 8-bit, 3 channels, BGR; alpha dropped, 16-bit samples reduced to their high byte
bool imread_color(const std::string &name, cv::Mat &bgr, std::string &err)
{
    kf::png::Image im;
    if (!kf::png::read(name, im, &err)) return false;
    bgr = cv::Mat(im.height, im.width, CV_8UC3);
    unsigned char *dst = bgr.ptr<unsigned char>();
    const size_t npx = (size_t)im.width * im.height;
    for (size_t i = 0; i < npx; ++i)
    {
        unsigned char s[4] = {0, 0, 0, 0};
        for (int c = 0; c < im.channels; ++c)
            s[c] = im.bit_depth == 16 ? (unsigned char)(im.data16[i * im.channels + c] >> 8) : im.data8[i * im.channels + c];
        const bool grey = im.channels <= 2;
        dst[3 * i + 0] = grey ? s[0] : s[2];
        dst[3 * i + 1] = grey ? s[0] : s[1];
        dst[3 * i + 2] = s[0];
    }
    return true;
}
// cv::imread(name, -1).convertTo(depth, CV_32FC1): the stored integers as floats (16-bit millimetres)
bool imread_depth(const std::string &name, cv::Mat &depth, std::string &err)
{
    kf::png::Image im;
    if (!kf::png::read(name, im, &err)) return false;
    if (im.channels != 1)
    {
        err = name + ": a depth image must have one channel";
        return false;
    }
    depth = cv::Mat(im.height, im.width, CV_32FC1);
    float *dst = depth.ptr<float>();
    const size_t npx = (size_t)im.width * im.height;
    for (size_t i = 0; i < npx; ++i) dst[i] = im.bit_depth == 16 ? (float)im.data16[i] : (float)im.data8[i];
    return true;
}
} // namespace

bool depth_sensor::open(const std::string &path)
{
    release();
    data_path = path;
    img_col_name = list_png(data_path + "/color");
    img_dep_name = list_png(data_path + "/depth");
    if (img_col_name.empty() || img_dep_name.empty())
    {
        err = "error: no camera!"; // depth_sensor.cpp:19 (the reference exits here)
        img_col_name.clear(); img_dep_name.clear();
        return false;
    }
    // intr.txt: up to nine numbers, those > 0.1 are kept and must be exactly fx cx fy cy scale (depth_sensor.cpp:23-46)
    kf::file::readIntrinsics(data_path + "/intr.txt", params);
    // image size from the first colour image (the reference reads it only when intr.txt was valid and otherwise
    // leaves 0 x 0; a usable size is returned here in both cases)
    kf::png::Image first;
    if (!kf::png::read(img_col_name[0], first, &err)) return false;
    params.width = first.width;
    params.height = first.height;
    return true;
}

bool depth_sensor::getFrame()
{
    if (img_dep_name.empty() || img_col_name.empty()) return false; // depth_sensor.cpp:187
    const bool ok = imread_color(img_col_name[0], color_map, err) && imread_depth(img_dep_name[0], depth_map, err);
    img_col_name.erase(img_col_name.begin());
    img_dep_name.erase(img_dep_name.begin());
    return ok;
}

void depth_sensor::release()
{
    img_col_name.clear();
    img_dep_name.clear();
}
