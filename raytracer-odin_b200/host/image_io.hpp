// image_io.hpp — texture decoding and image writing for the C++ host.
//   load_texture  = load_texture (textures.odin:25-68), i.e. stb_image semantics: Radiance .hdr decodes
//                   to 3 x f32, everything else to u8 with the file's native channel count
//                   (16-bit PNG samples keep their high byte, like stbi's 16 -> 8 conversion).
//   write_png/ppm = the two formats save_result writes (output.odin:82-107).
// Decoders: PNG (non-interlaced, all colour types / bit depths, zlib) and Radiance RGBE (flat and
// new-style RLE).  JPEG & co. are not decoded: the synthetic BASELINE scenes use PNG + .hdr only.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace orh {

struct Texture { // Texture (textures.odin:14-19)
    std::vector<uint8_t> u8;
    std::vector<float> f32;
    int width = 0, height = 0, channels = 0;
    bool is_f32 = false;
    const void* data() const { return is_f32 ? (const void*)f32.data() : (const void*)u8.data(); }
};

// Returns false with *err = the reference's message ("Failed to read texture file: ..." /
// "Failed to parse texture", textures.odin:28,56) or a more specific reason.
bool load_texture(const std::string& path, Texture* out, std::string* err);
bool decode_png(const uint8_t* data, size_t size, Texture* out, std::string* err);
bool decode_hdr(const uint8_t* data, size_t size, Texture* out, std::string* err);

bool write_png(const std::string& path, int w, int h, const uint8_t* rgb, std::string* err);
bool write_ppm(const std::string& path, int w, int h, const uint8_t* rgb, std::string* err);

bool read_file(const std::string& path, std::vector<uint8_t>* out);

} // namespace orh
