// x3_cli.cpp -- the `x3` command line tool of the reference (src/bin/x3.rs:33-82): direction by extension.
//   x3 -i in.wav -o out.x3a      encode        x3 -i in.x3a -o out.wav      decode
#include <cstdio>
#include <cstring>
#include <string>

#include "x3.hpp"

static bool ends_with(const std::string &s, const char *suf) {
  size_t n = std::strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

int main(int argc, char **argv) {
  std::string in, out;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if ((a == "-i" || a == "--input") && i + 1 < argc) in = argv[++i];
    else if ((a == "-o" || a == "--output") && i + 1 < argc) out = argv[++i];
    else { std::fprintf(stderr, "usage: x3 -i <in.wav|in.x3a> -o <out.x3a|out.wav>\n"); return 2; }
  }
  if (in.empty() || out.empty()) { std::fprintf(stderr, "usage: x3 -i <in.wav|in.x3a> -o <out.x3a|out.wav>\n"); return 2; }
  try {
    if (ends_with(in, ".wav") && ends_with(out, ".x3a")) x3::encodefile::wav_to_x3a(in, out);
    else if (ends_with(in, ".x3a") && ends_with(out, ".wav")) x3::decodefile::x3a_to_wav(in, out);
    else { std::fprintf(stderr, "Invalid audio file, expecting a '.wav' or '.x3a' file\n"); return 2; }  // bin/x3.rs:35-40
  } catch (const x3::X3Error &e) {
    std::fprintf(stderr, "X3Error(%d): %s\n", e.code(), e.what());
    return 1;
  }
  return 0;
}
