// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <boost/property_tree/json_parser.hpp>.
// read_json builds the same tree shape Boost does: objects -> keyed children, arrays -> children
// with empty keys, scalars -> data strings (true/false/null kept as the literal text).
#ifndef CHIML_ORACLE_SHIM_JSON_PARSER_HPP
#define CHIML_ORACLE_SHIM_JSON_PARSER_HPP

#include <boost/property_tree/ptree.hpp>
#include <cctype>
#include <fstream>
#include <iterator>

namespace boost { namespace property_tree { namespace json_parser {

class json_parser_error : public ptree_error
{ public: explicit json_parser_error(const std::string& w) : ptree_error(w) {} };

namespace detail {
struct Reader
{
    const std::string& s;
    std::size_t i;
    explicit Reader(const std::string& str) : s(str), i(0) {}
    void ws() { while(i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) ++i; }
    [[noreturn]] void fail(const std::string& what) { throw json_parser_error("json: " + what + " at offset " + std::to_string(i)); }
    std::string str()
    {
        if(s[i] != '"') fail("expected string");
        ++i;
        std::string out;
        while(i < s.size() && s[i] != '"')
        {
            if(s[i] == '\\')
            {
                ++i;
                if(i >= s.size()) fail("bad escape");
                switch(s[i])
                {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u':
                    {
                        if(i + 4 >= s.size()) fail("bad \\u escape");
                        unsigned cp = std::stoul(s.substr(i + 1, 4), nullptr, 16);
                        i += 4;
                        if(cp < 0x80) out += char(cp);
                        else if(cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
                        else { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += s[i];
                }
                ++i;
            }
            else
                out += s[i++];
        }
        if(i >= s.size()) fail("unterminated string");
        ++i;
        return out;
    }
    void value(ptree& node)
    {
        ws();
        if(i >= s.size()) fail("unexpected end");
        char c = s[i];
        if(c == '{')
        {
            ++i; ws();
            if(s[i] == '}') { ++i; return; }
            while(true)
            {
                ws();
                std::string key = str();
                ws();
                if(s[i] != ':') fail("expected ':'");
                ++i;
                ptree child;
                value(child);
                node.push_back(ptree::value_type(key, child));
                ws();
                if(s[i] == ',') { ++i; continue; }
                if(s[i] == '}') { ++i; return; }
                fail("expected ',' or '}'");
            }
        }
        else if(c == '[')
        {
            ++i; ws();
            if(s[i] == ']') { ++i; return; }
            while(true)
            {
                ptree child;
                value(child);
                node.push_back(ptree::value_type("", child));
                ws();
                if(s[i] == ',') { ++i; continue; }
                if(s[i] == ']') { ++i; return; }
                fail("expected ',' or ']'");
            }
        }
        else if(c == '"')
            node.data() = str();
        else
        {
            std::size_t b = i;
            while(i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']' && !std::isspace(static_cast<unsigned char>(s[i]))) ++i;
            if(i == b) fail("expected value");
            node.data() = s.substr(b, i - b);
        }
    }
};
} // namespace detail

inline void read_json(std::istream& in, ptree& pt)
{
    std::string text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    detail::Reader r(text);
    ptree root;
    r.value(root);
    r.ws();
    if(r.i != text.size()) r.fail("trailing characters");
    pt = root;
}

inline void read_json(const std::string& filename, ptree& pt)
{
    std::ifstream in(filename.c_str());
    if(!in) throw json_parser_error("cannot open file " + filename);
    read_json(in, pt);
}

} // namespace json_parser

using json_parser::read_json;

}} // namespace boost::property_tree

#endif
