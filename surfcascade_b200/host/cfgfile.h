// Reader / writer for the libconfig 1.4.9 text format the reference stores its model in (model.cfg).
//
// The reference links the vendored libconfig (libconfig/libconfig.c, scanner.c, grammar.c) through
// libconfig++ (Model.cpp:21-193).  This is an independent implementation of the subset of that format a
// model file uses -- groups { }, lists ( ), arrays [ ], `name = value;` / `name : value;`, ints, int64 (L suffix),
// hex, floats, booleans, strings, # // and /* */ comments -- with the two behaviours Model relies on:
//   * no automatic int<->float conversion on lookup (libconfigcpp.c++:1137-1145 assertType), and
//   * floats written with "%.10g", forced to contain '.' or 'e' (libconfig.c:212-243, FLOAT_PRECISION 10).
#ifndef SC_CFGFILE_H
#define SC_CFGFILE_H

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace sccfg {

struct ParseError : std::runtime_error {
    int line;
    ParseError(const std::string& what, int line_) : std::runtime_error(what), line(line_) {}
};
struct NotFound : std::runtime_error {
    explicit NotFound(const std::string& path) : std::runtime_error("setting not found: " + path) {}
};
struct TypeMismatch : std::runtime_error {
    explicit TypeMismatch(const std::string& path) : std::runtime_error("setting type mismatch: " + path) {}
};
struct IoError : std::runtime_error {
    explicit IoError(const std::string& what) : std::runtime_error(what) {}
};

class Node {
public:
    enum Kind { Group, List, Array, Int, Int64, Float, Bool, String };

    explicit Node(Kind k = Group) : kind_(k), i_(0), f_(0.0), b_(false) {}

    Kind kind() const { return kind_; }
    int length() const { return (int)kids_.size(); }

    // lookup; throw NotFound like libconfig's Setting::operator[]
    Node& operator[](const char* name);
    const Node& operator[](const char* name) const;
    Node& operator[](int index);
    const Node& operator[](int index) const;
    bool exists(const char* name) const;

    // typed reads; throw TypeMismatch unless the stored type matches (ints widen to int64 only)
    int asInt() const;
    long long asInt64() const;
    double asDouble() const;
    float asFloat() const { return (float)asDouble(); }  // (float) config_setting_get_float, libconfigcpp.c++:710-716
    bool asBool() const;
    const std::string& asString() const;

    // building
    Node& add(const std::string& name, Kind k);  // child of a group
    Node& add(Kind k);                           // element of a list / array
    Node& set(int v) { kind_ = Int; i_ = v; return *this; }
    Node& set(long long v) { kind_ = Int64; i_ = v; return *this; }
    Node& set(double v) { kind_ = Float; f_ = v; return *this; }
    Node& set(bool v) { kind_ = Bool; b_ = v; return *this; }
    Node& set(const std::string& v) { kind_ = String; s_ = v; return *this; }

    const std::string& name() const { return name_; }

private:
    friend class File;
    Kind kind_;
    std::string name_;
    long long i_;
    double f_;
    bool b_;
    std::string s_;
    std::vector<std::unique_ptr<Node>> kids_;
};

class File {
public:
    File() : root_(Node::Group) {}
    Node& root() { return root_; }
    void read(const std::string& path);        // IoError / ParseError
    void parse(const std::string& text);       // ParseError
    void write(const std::string& path) const; // IoError
    std::string str() const;

private:
    Node root_;
};

// "%.10g" with a forced ".0" when the text would read back as an integer (libconfig.c:212-243).
std::string format_float(double v);

}  // namespace sccfg

#endif
