// Options.hpp -- CLI contract of the reference (include/Options.hpp:13-124):
//   -f/-F file, -k/-K K (32), -a/-A alpha (0.3), -d/-D delta (0.3), -t/-T test sweep, -l/-L log dir;
//   with no option flags: argv[1] = file, argv[2] = K.
// Additions (absent from the reference): -b block_size (0 = reference rule), -g free-memory bytes used
// by that rule (reproducibility), -i iterations; multi-GPU (one process per GPU):
// -r rank, -w world size, -u path of the file through which rank 0 hands the 128-byte NCCL id to the others.
#pragma once
#include <cstdlib>
#include <iostream>
#include <string>
#include <unordered_map>

class Options {
 public:
  Options(int argc, const char* const argv[]) {
    const std::string self = argv[0];
    const size_t slash = self.find_last_of('/');
    programPath_ = slash == std::string::npos ? "" : self.substr(0, slash + 1);
    programName_ = slash == std::string::npos ? self : self.substr(slash + 1);
    std::unordered_map<std::string, std::string> opts;
    for (int i = 1; i < argc; ++i) {
      if (argv[i][0] != '-') continue;
      const std::string o = argv[i];
      if (opts.count(o)) { std::cerr << "Option " << o << "is duplicated." << std::endl; continue; }
      if (i + 1 >= argc) { std::cerr << "Option " << o << "requires an argument." << std::endl; continue; }
      opts[o] = argv[i + 1];
    }
    for (const auto& kv : opts) parse(kv.first, kv.second);
    if (opts.empty() && argc > 1) {
      inputFile_ = argv[1];
      if (argc > 2) K_ = std::stoi(argv[2]);
    }
  }
  std::string programPath() const { return programPath_; }
  std::string programName() const { return programName_; }
  std::string inputFile() const { return inputFile_; }
  size_t K() const { return K_; }
  int numIterations() const { return numIterations_; }
  float similarityThresholdAlpha() const { return alpha_; }
  float blockDensityThresholdDelta() const { return delta_; }
  bool testMode() const { return testMode_; }
  std::string outputLogDirectory() const { return logDir_; }
  uint32_t blockSize() const { return blockSize_; }
  uint64_t freeMemForBlockSize() const { return freeMem_; }
  int rank() const { return rank_; }
  int world() const { return world_; }
  std::string idFile() const { return idFile_; }

 private:
  void parse(const std::string& o, const std::string& v) {
    try {
      if (o == "-F" || o == "-f") inputFile_ = v;
      if (o == "-K" || o == "-k") K_ = std::stoi(v);
      if (o == "-A" || o == "-a") alpha_ = std::stof(v);
      if (o == "-D" || o == "-d") delta_ = std::stof(v);
      if (o == "-T" || o == "-t") testMode_ = std::stoi(v);
      if (o == "-L" || o == "-l") logDir_ = v;
      if (o == "-B" || o == "-b") blockSize_ = static_cast<uint32_t>(std::stoul(v));
      if (o == "-G" || o == "-g") freeMem_ = std::stoull(v);
      if (o == "-I" || o == "-i") numIterations_ = std::stoi(v);
      if (o == "-R" || o == "-r") rank_ = std::stoi(v);
      if (o == "-W" || o == "-w") world_ = std::stoi(v);
      if (o == "-U" || o == "-u") idFile_ = v;
    } catch (const std::exception& e) {
      std::cerr << "Invalid argument: " << e.what() << std::endl;
    }
  }
  std::string programPath_, programName_, inputFile_, logDir_;
  size_t K_ = 32;
  int numIterations_ = 10;
  float alpha_ = 0.3f, delta_ = 0.3f;
  bool testMode_ = false;
  uint32_t blockSize_ = 0;
  uint64_t freeMem_ = 0;
  int rank_ = 0, world_ = 1;
  std::string idFile_;
};
