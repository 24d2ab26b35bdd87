// Minimal JSON reader for chiML input files (host side of the B200 engine).
// The reference parses its inputs with Boost.PropertyTree (INPUTS/parallelInputs.cpp:12-840): every
// scalar is kept as text and converted on access, objects keep insertion order, arrays are children
// with empty keys, `//` comments are stripped first (stripComments, parallelInputs.cpp:1625-1649).
// This tree offers the same access semantics (get<T>(path), get<T>(path, default), child lists)
// so the input contract -- including "missing key -> default" and "unparsable value -> default"
// -- is the same.
#pragma once

#include <cctype>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace chiml_host {

class Json
{
public:
    std::string data;                                   // scalar text ("" for containers)
    std::vector<std::pair<std::string, Json>> kids;     // ordered children (key "" inside arrays)

    size_t size() const { return kids.size(); }

    const Json* find(const std::string& path) const
    {
        const Json* cur = this;
        size_t start = 0;
        if(path.empty()) return cur;
        while(true)
        {
            const size_t dot = path.find('.', start);
            const std::string key = path.substr(start, dot == std::string::npos ? std::string::npos : dot - start);
            const Json* next = nullptr;
            for(const auto& kv : cur->kids)
                if(kv.first == key) { next = &kv.second; break; }
            if(!next) return nullptr;
            cur = next;
            if(dot == std::string::npos) return cur;
            start = dot + 1;
        }
    }
    const Json& child(const std::string& path) const
    {
        const Json* p = find(path);
        if(!p) throw std::runtime_error("input file: no such node (" + path + ")");
        return *p;
    }
    template <typename T> static bool convert(const std::string& s, T& out)
    {
        std::istringstream iss(s);
        iss >> out;
        if(iss.fail()) return false;
        iss >> std::ws;
        return iss.eof();
    }
    template <typename T> T value() const
    {
        T out;
        if(!convertT(data, out)) throw std::runtime_error("input file: cannot convert \"" + data + "\"");
        return out;
    }
    template <typename T> T get(const std::string& path) const { return child(path).template value<T>(); }
    template <typename T> T get(const std::string& path, const T& dflt) const
    {
        const Json* p = find(path);
        if(!p) return dflt;
        T out;
        return convertT(p->data, out) ? out : dflt;
    }
    std::string get(const std::string& path, const char* dflt) const { return get<std::string>(path, std::string(dflt)); }

private:
    template <typename T> static bool convertT(const std::string& s, T& out) { return convert<T>(s, out); }
};

template <> inline bool Json::convert<std::string>(const std::string& s, std::string& out) { out = s; return true; }
template <> inline bool Json::convert<bool>(const std::string& s, bool& out)
{
    // Boost's stream translator: numeric 0/1 first, then boolalpha
    if(s == "true") { out = true; return true; }
    if(s == "false") { out = false; return true; }
    if(s == "1") { out = true; return true; }
    if(s == "0") { out = false; return true; }
    return false;
}

namespace detail {
struct JsonReader
{
    const std::string& s;
    size_t i = 0;
    explicit JsonReader(const std::string& str) : s(str) {}
    void ws() { while(i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i; }
    [[noreturn]] void fail(const std::string& what) const { throw std::runtime_error("input file: JSON " + what + " at offset " + std::to_string(i)); }
    std::string str()
    {
        if(i >= s.size() || s[i] != '"') fail("expected string");
        ++i;
        std::string out;
        while(i < s.size() && s[i] != '"')
        {
            if(s[i] == '\\' && i + 1 < s.size())
            {
                ++i;
                switch(s[i]) { case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; default: out += s[i]; }
                ++i;
            }
            else out += s[i++];
        }
        if(i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
    void value(Json& node)
    {
        ws();
        if(i >= s.size()) fail("unexpected end");
        const char c = s[i];
        if(c == '{')
        {
            ++i; ws();
            if(i < s.size() && s[i] == '}') { ++i; return; }
            while(true)
            {
                ws();
                std::string key = str();
                ws();
                if(i >= s.size() || s[i] != ':') fail("expected ':'");
                ++i;
                node.kids.emplace_back(key, Json());
                value(node.kids.back().second);
                ws();
                if(i < s.size() && s[i] == ',') { ++i; continue; }
                if(i < s.size() && s[i] == '}') { ++i; return; }
                fail("expected ',' or '}'");
            }
        }
        else if(c == '[')
        {
            ++i; ws();
            if(i < s.size() && s[i] == ']') { ++i; return; }
            while(true)
            {
                node.kids.emplace_back(std::string(), Json());
                value(node.kids.back().second);
                ws();
                if(i < s.size() && s[i] == ',') { ++i; continue; }
                if(i < s.size() && s[i] == ']') { ++i; return; }
                fail("expected ',' or ']'");
            }
        }
        else if(c == '"') node.data = str();
        else
        {
            const size_t b = i;
            while(i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']' && !std::isspace(static_cast<unsigned char>(s[i]))) ++i;
            if(i == b) fail("expected value");
            node.data = s.substr(b, i - b);
        }
    }
};
} // namespace detail

// parse a chiML input file; `//` comments are removed line by line exactly as stripComments does
inline Json read_input_file(const std::string& filename)
{
    std::ifstream in(filename.c_str());
    if(!in) throw std::runtime_error("cannot open input file " + filename);
    std::string text, line;
    while(std::getline(in, line))
    {
        const size_t f1 = line.find('/');
        const size_t f2 = f1 == std::string::npos ? std::string::npos : line.find('/', f1 + 1);
        if(f1 != std::string::npos && f2 == f1 + 1) line.erase(f1);
        text += line;
        text += '\n';
    }
    detail::JsonReader r(text);
    Json root;
    r.value(root);
    r.ws();
    if(r.i != text.size()) r.fail("trailing characters");
    return root;
}

} // namespace chiml_host
