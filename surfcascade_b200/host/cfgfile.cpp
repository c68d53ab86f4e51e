// See cfgfile.h.  Recursive-descent reader and pretty-printer for libconfig-format model files.
#include "cfgfile.h"

#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace sccfg {

// ---- Node ------------------------------------------------------------------------------------------------
Node& Node::operator[](const char* name) {
    for (auto& k : kids_)
        if (k->name_ == name) return *k;
    throw NotFound(name);
}
const Node& Node::operator[](const char* name) const { return const_cast<Node*>(this)->operator[](name); }
Node& Node::operator[](int index) {
    if (index < 0 || index >= (int)kids_.size()) throw NotFound("[" + std::to_string(index) + "]");
    return *kids_[index];
}
const Node& Node::operator[](int index) const { return const_cast<Node*>(this)->operator[](index); }
bool Node::exists(const char* name) const {
    for (auto& k : kids_)
        if (k->name_ == name) return true;
    return false;
}
int Node::asInt() const {
    if (kind_ != Int) throw TypeMismatch(name_);
    return (int)i_;
}
long long Node::asInt64() const {
    if (kind_ != Int && kind_ != Int64) throw TypeMismatch(name_);
    return i_;
}
double Node::asDouble() const {
    if (kind_ != Float) throw TypeMismatch(name_);
    return f_;
}
bool Node::asBool() const {
    if (kind_ != Bool) throw TypeMismatch(name_);
    return b_;
}
const std::string& Node::asString() const {
    if (kind_ != String) throw TypeMismatch(name_);
    return s_;
}
Node& Node::add(const std::string& name, Kind k) {
    kids_.emplace_back(new Node(k));
    kids_.back()->name_ = name;
    return *kids_.back();
}
Node& Node::add(Kind k) {
    kids_.emplace_back(new Node(k));
    return *kids_.back();
}

// ---- reader ----------------------------------------------------------------------------------------------
namespace {

struct Lexer {
    const std::string& t;
    size_t i = 0;
    int line = 1;
    explicit Lexer(const std::string& text) : t(text) {}

    [[noreturn]] void die(const std::string& what) const { throw ParseError(what, line); }

    void skip() {
        for (;;) {
            while (i < t.size() && isspace((unsigned char)t[i])) { if (t[i] == '\n') line++; i++; }
            if (i >= t.size()) return;
            if (t[i] == '#' || (t[i] == '/' && i + 1 < t.size() && t[i + 1] == '/')) {
                while (i < t.size() && t[i] != '\n') i++;
            } else if (t[i] == '/' && i + 1 < t.size() && t[i + 1] == '*') {
                i += 2;
                while (i + 1 < t.size() && !(t[i] == '*' && t[i + 1] == '/')) { if (t[i] == '\n') line++; i++; }
                if (i + 1 >= t.size()) die("unterminated comment");
                i += 2;
            } else
                return;
        }
    }
    bool eof() { skip(); return i >= t.size(); }
    char peek() { skip(); return i < t.size() ? t[i] : '\0'; }
    bool take(char c) { if (peek() == c) { i++; return true; } return false; }
    void expect(char c) { if (!take(c)) die(std::string("expected '") + c + "'"); }

    std::string name() {
        skip();
        size_t b = i;
        if (i < t.size() && (isalpha((unsigned char)t[i]) || t[i] == '*')) {
            i++;
            while (i < t.size() && (isalnum((unsigned char)t[i]) || t[i] == '_' || t[i] == '-' || t[i] == '*')) i++;
        }
        if (b == i) die("expected a setting name");
        return t.substr(b, i - b);
    }
};

void parse_scalar(Lexer& lx, Node& n) {
    lx.skip();
    const std::string& t = lx.t;
    size_t b = lx.i;
    if (t[b] == '"') {
        std::string s;
        // adjacent string literals concatenate, like the libconfig scanner
        while (lx.peek() == '"') {
            lx.i++;
            while (lx.i < t.size() && t[lx.i] != '"') {
                char c = t[lx.i++];
                if (c == '\\' && lx.i < t.size()) {
                    char e = t[lx.i++];
                    switch (e) {
                        case 'n': s += '\n'; break;
                        case 'r': s += '\r'; break;
                        case 't': s += '\t'; break;
                        case 'f': s += '\f'; break;
                        case 'x': {
                            if (lx.i + 1 >= t.size()) lx.die("bad \\x escape");
                            s += (char)strtol(t.substr(lx.i, 2).c_str(), nullptr, 16);
                            lx.i += 2;
                            break;
                        }
                        default: s += e;
                    }
                } else {
                    if (c == '\n') lx.line++;
                    s += c;
                }
            }
            if (lx.i >= t.size()) lx.die("unterminated string");
            lx.i++;
        }
        n.set(s);
        return;
    }
    size_t e = b;
    while (e < t.size() && (isalnum((unsigned char)t[e]) || t[e] == '+' || t[e] == '-' || t[e] == '.')) e++;
    if (e == b) lx.die("expected a value");
    std::string tok = t.substr(b, e - b);
    lx.i = e;
    std::string low;
    for (char c : tok) low += (char)tolower((unsigned char)c);
    if (low == "true") { n.set(true); return; }
    if (low == "false") { n.set(false); return; }
    const bool hex = low.size() > 2 && low[0] == '0' && low[1] == 'x';
    bool is_float = false;
    if (!hex)
        for (char c : low)
            if (c == '.' || c == 'e') is_float = true;
    if (is_float) {
        // libconfig's scanner hands the token to atof(), libconfig/scanner.c:1146
        char* end = nullptr;
        double v = strtod(tok.c_str(), &end);
        if (!end || *end) lx.die("bad float '" + tok + "'");
        n.set(v);
        return;
    }
    bool wide = false;
    while (!low.empty() && low.back() == 'l') { low.pop_back(); wide = true; }
    if (low.empty()) lx.die("bad integer '" + tok + "'");
    char* end = nullptr;
    errno = 0;
    long long v = hex ? (long long)strtoull(low.c_str(), &end, 16) : strtoll(low.c_str(), &end, 10);
    if (!end || *end || errno) lx.die("bad integer '" + tok + "'");
    if (wide || v > 2147483647LL || v < -2147483648LL) n.set(v);
    else n.set((int)v);
}

}  // namespace

void File::parse(const std::string& text) {
    Node fresh(Node::Group);
    Lexer lx(text);
    struct Rec {
        static void settings(Lexer& lx, Node& group, char closer) {
            for (;;) {
                if (closer ? lx.peek() == closer : lx.eof()) return;
                if (lx.eof()) lx.die("unexpected end of file");
                std::string nm = lx.name();
                if (!lx.take('=') && !lx.take(':')) lx.die("expected '=' or ':'");
                Node& child = group.add(nm, Node::Group);
                value(lx, child);
                child.name_ = nm;
                if (!lx.take(';')) lx.take(',');
            }
        }
        static void value(Lexer& lx, Node& n) {
            char c = lx.peek();
            if (c == '{') {
                lx.i++;
                n.kind_ = Node::Group;
                settings(lx, n, '}');
                lx.expect('}');
            } else if (c == '(' || c == '[') {
                const char closer = c == '(' ? ')' : ']';
                lx.i++;
                n.kind_ = c == '(' ? Node::List : Node::Array;
                while (lx.peek() != closer) {
                    if (lx.eof()) lx.die("unexpected end of file in list");
                    Node& el = n.add(Node::Group);
                    value(lx, el);
                    if (n.kind_ == Node::Array && (el.kind_ == Node::Group || el.kind_ == Node::List || el.kind_ == Node::Array))
                        lx.die("arrays hold scalars only");
                    lx.take(',');
                }
                lx.expect(closer);
            } else {
                parse_scalar(lx, n);
            }
        }
    };
    Rec::settings(lx, fresh, '\0');
    root_ = std::move(fresh);
}

void File::read(const std::string& path) {
    std::ifstream in(path.c_str(), std::ios::binary);
    if (!in.good()) throw IoError("cannot open " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    parse(ss.str());
}

// ---- writer ----------------------------------------------------------------------------------------------
std::string format_float(double v) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%.10g", v);
    if (!strchr(buf, 'e') && !strchr(buf, '.') && !strchr(buf, 'n') && !strchr(buf, 'i')) strcat(buf, ".0");
    return buf;
}

std::string File::str() const {
    std::string out;
    struct W {
        static void value(std::string& out, const Node& n, int depth) {
            switch (n.kind()) {
                case Node::Int: out += std::to_string((int)n.i_); break;
                case Node::Int64: out += std::to_string(n.i_) + "L"; break;
                case Node::Float: out += format_float(n.f_); break;
                case Node::Bool: out += n.b_ ? "true" : "false"; break;
                case Node::String: {
                    out += '"';
                    for (char ch : n.s_) {
                        if (ch == '"' || ch == '\\') { out += '\\'; out += ch; }
                        else if (ch == '\n') out += "\\n";
                        else if (ch == '\r') out += "\\r";
                        else if (ch == '\t') out += "\\t";
                        else if (ch == '\f') out += "\\f";
                        else out += ch;
                    }
                    out += '"';
                    break;
                }
                case Node::Group:
                    out += "\n";
                    out.append((size_t)depth * 2, ' ');
                    out += "{\n";
                    settings(out, n, depth + 1);
                    out.append((size_t)depth * 2, ' ');
                    out += "}";
                    break;
                case Node::List:
                case Node::Array: {
                    const bool list = n.kind() == Node::List;
                    out += list ? "( " : "[ ";
                    for (size_t i = 0; i < n.kids_.size(); i++) {
                        value(out, *n.kids_[i], depth + 1);
                        if (i + 1 < n.kids_.size()) out += ",";
                        out += " ";
                    }
                    out += list ? ")" : "]";
                    break;
                }
            }
        }
        static void settings(std::string& out, const Node& g, int depth) {
            for (auto& k : g.kids_) {
                out.append((size_t)depth * 2, ' ');
                out += k->name_;
                out += (k->kind_ == Node::Group) ? " : " : " = ";
                value(out, *k, depth);
                out += ";\n";
            }
        }
    };
    W::settings(out, root_, 0);
    return out;
}

void File::write(const std::string& path) const {
    std::ofstream of(path.c_str(), std::ios::binary);
    if (!of.good()) throw IoError("cannot open " + path + " for writing");
    of << str();
    if (!of.good()) throw IoError("write failed: " + path);
}

}  // namespace sccfg
