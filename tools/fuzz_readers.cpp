// fuzz_readers.cpp — sanitizer harness for the host-side file readers (MJCF / XML, the binary model format, STL / OBJ
// meshes, PNG and binary height fields).  These take files from outside the process (mujoco_env.cpp:771-911 loads
// whatever path or string the caller names), so malformed input has to end in an error code, never in a crash.
//   g++ -std=c++17 -g -O1 -fsanitize=address,undefined -fno-sanitize-recover=undefined -Iinclude \
//       -Imujoco_ros_pkgs_b200/csrc tools/fuzz_readers.cpp mujoco_ros_pkgs_b200/csrc/model/*.cpp -o /tmp/fuzz_readers
//   /tmp/fuzz_readers <dir-with-seed-files> <iterations> [seed]
// Every seed file is mutated (bit flips, byte splices, truncation, duplicated ranges, number edits) and fed through the
// reader its extension selects; *.xml seeds may reference sibling asset files, which are mutated in their turn.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "b2mj.h"

static std::string slurp(const std::string& p) {
  std::ifstream f(p, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
static void spit(const std::string& p, const std::string& s) {
  std::ofstream f(p, std::ios::binary);
  f.write(s.data(), (std::streamsize)s.size());
}
static bool ends(const std::string& s, const char* e) {
  const size_t n = std::strlen(e);
  return s.size() >= n && s.compare(s.size() - n, n, e) == 0;
}

static std::string mutate(std::string s, std::mt19937& rng, bool text) {
  if (s.empty()) return s;
  const int edits = 1 + (int)(rng() % 4);
  for (int e = 0; e < edits && !s.empty(); e++) {
    const size_t i = rng() % s.size();
    switch (rng() % 7) {
      case 0: s[i] = (char)(s[i] ^ (1 << (rng() % 8))); break;
      case 1: s[i] = text ? "<>\"'/= -0123456789.e\n"[rng() % 21] : (char)(rng() & 0xFF); break;
      case 2: s.resize(i); break;
      case 3: { const size_t n = 1 + rng() % 64; s.erase(i, n); break; }
      case 4: { const size_t n = 1 + rng() % 64, j = rng() % s.size(); s.insert(i, s.substr(j, n)); break; }
      case 5: {  // overwrite 4 bytes with an extreme integer (binary headers: counts, sizes)
        static const uint32_t v[] = {0u, 1u, 0x7FFFFFFFu, 0x80000000u, 0xFFFFFFFFu, 0x10000u, 0xFFFFu};
        const uint32_t x = v[rng() % 7];
        for (int k = 0; k < 4 && i + k < s.size(); k++) s[i + k] = (char)(x >> (8 * k));
        break;
      }
      default: {  // text: replace a number by an extreme one
        static const char* v[] = {"1e308", "-1e308", "nan", "0", "-1", "1e-320", "99999999999", ""};
        size_t a = s.find_first_of("0123456789", i);
        if (text && a != std::string::npos) {
          size_t b = s.find_first_not_of("0123456789.e-+", a);
          s.replace(a, b == std::string::npos ? std::string::npos : b - a, v[rng() % 8]);
        }
      }
    }
  }
  return s;
}

int main(int argc, char** argv) {
  if (argc < 3) return std::fprintf(stderr, "usage: %s seed_dir iterations [seed]\n", argv[0]), 2;
  const std::string dir = argv[1];
  const long iters = std::atol(argv[2]);
  std::mt19937 rng(argc > 3 ? (unsigned)std::atol(argv[3]) : 1u);
  std::vector<std::string> names;
  if (DIR* d = opendir(dir.c_str())) {
    while (dirent* e = readdir(d))
      if (e->d_name[0] != '.' && !std::strstr(e->d_name, "_fz")) names.push_back(e->d_name);
    closedir(d);
  }
  if (names.empty()) return std::fprintf(stderr, "no seed files in %s\n", dir.c_str()), 2;
  std::vector<std::string> seeds;
  for (auto& n : names) seeds.push_back(slurp(dir + "/" + n));
  long ok = 0, refused = 0;
  for (long it = 0; it < iters; it++) {
    const size_t k = rng() % names.size();
    const std::string& name = names[k];
    b2mjModel* m = nullptr;
    int rc;
    if (ends(name, ".xml")) {
      // mutate the XML itself or one of the asset files it may name (restored afterwards)
      if (rng() % 2 == 0) {
        size_t a = rng() % names.size();
        if (!ends(names[a], ".xml")) {
          // the scene that names every asset seed, when the corpus has it (tools/fuzz_seeds.py)
          std::string user = name;
          for (auto& nm : names) if (nm == "assets_scene.xml") user = nm;
          spit(dir + "/" + names[a], mutate(seeds[a], rng, ends(names[a], ".obj") || ends(names[a], "_ascii.stl")));
          rc = b2mj_model_from_xml_file((dir + "/" + user).c_str(), &m);
          spit(dir + "/" + names[a], seeds[a]);
        } else {
          rc = b2mj_model_from_xml_file((dir + "/" + name).c_str(), &m);
        }
      } else {
        const std::string p = dir + "/" + name.substr(0, name.size() - 4) + "_fz.xml";
        spit(p, mutate(seeds[k], rng, true));
        rc = b2mj_model_from_xml_file(p.c_str(), &m);
      }
    } else if (ends(name, ".b2mjb")) {
      const std::string p = dir + "/cur_fz.b2mjb";
      spit(p, mutate(seeds[k], rng, false));
      rc = b2mj_model_load_binary(p.c_str(), &m);
    } else {
      continue;  // asset seeds are reached through the XML that names them
    }
    if (rc == 0 && m) {
      ok++;
      // a model that loaded must survive a save / load round trip and set_const
      const std::string p = dir + "/rt_fz.b2mjb";
      if (b2mj_model_save_binary(m, p.c_str()) == 0) {
        b2mjModel* m2 = nullptr;
        if (b2mj_model_load_binary(p.c_str(), &m2) == 0 && m2) b2mj_model_free(m2);
      }
      b2mj_model_free(m);
    } else {
      refused++;
      if (!b2mj_last_error() || !*b2mj_last_error()) { std::fprintf(stderr, "iteration %ld: failure without a message\n", it); return 1; }
    }
  }
  std::printf("fuzz_readers: %ld inputs, %ld loaded, %ld refused with a message, 0 crashes\n", iters, ok, refused);
  return 0;
}
