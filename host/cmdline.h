// cmdline.h - a small flag parser with the surface the reference command lines get from tclap
// (EncodeParams.cpp:52-130, DecodeParams.cpp:30-80): "-x 1920" / "--width 1920" value flags, "-v" switches,
// two positional file names, errors reported as "Command line error: ...".
#ifndef VC2_CMDLINE_H
#define VC2_CMDLINE_H
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace vc2cli {

struct Flag { std::string shortName, longName; bool isSwitch; };

class CmdLine {
 public:
  void add(const std::string& s, const std::string& l, bool isSwitch) { flags_.push_back(Flag{s, l, isSwitch}); }
  void parse(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
      const std::string a = argv[i];
      if (a.size() >= 2 && a[0] == '-' && a != "-") {
        const Flag* f = find(a);
        if (!f) throw std::invalid_argument("Command line error: Couldn't find match for argument for arg " + a);
        if (values_.count(f->shortName)) throw std::invalid_argument("Command line error: Argument already set! for arg -" + f->shortName);
        if (f->isSwitch) values_[f->shortName] = "1";
        else {
          if (i + 1 >= argc) throw std::invalid_argument("Command line error: Missing a value for this argument! for arg -" + f->shortName);
          values_[f->shortName] = argv[++i];
        }
      } else positional_.push_back(a);
    }
  }
  bool isSet(const std::string& s) const { return values_.count(s) != 0; }
  std::string str(const std::string& s, const std::string& dflt = "") const {
    std::map<std::string, std::string>::const_iterator it = values_.find(s);
    return it == values_.end() ? dflt : it->second;
  }
  int integer(const std::string& s, int dflt) const {
    if (!isSet(s)) return dflt;
    const std::string v = str(s);
    size_t pos = 0;
    int r = 0;
    try { r = std::stoi(v, &pos); } catch (...) { pos = 0; }
    if (pos != v.size()) throw std::invalid_argument("Command line error: Couldn't read argument value from string '" + v + "' for arg -" + s);
    return r;
  }
  const std::vector<std::string>& positional() const { return positional_; }
 private:
  const Flag* find(const std::string& a) const {
    for (size_t i = 0; i < flags_.size(); ++i)
      if (a == "-" + flags_[i].shortName || a == "--" + flags_[i].longName) return &flags_[i];
    return 0;
  }
  std::vector<Flag> flags_;
  std::map<std::string, std::string> values_;
  std::vector<std::string> positional_;
};

}  // namespace vc2cli
#endif
