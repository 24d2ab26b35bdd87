// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <boost/property_tree/ptree.hpp> so that the
// reference's INPUTS/ translation unit compiles in place without Boost.  Provides the subset of
// basic_ptree<std::string,std::string> the reference touches: ordered children, get<T>(path[,default]),
// get_child(path), get_value<T>(), size(), iteration, put/add_child/push_back.  Type conversion
// follows Boost's stream translator (whole string must be consumed; bool accepts true/false/1/0).
#ifndef CHIML_ORACLE_SHIM_PTREE_HPP
#define CHIML_ORACLE_SHIM_PTREE_HPP

#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <ios>

namespace boost { namespace property_tree {

class ptree_error : public std::runtime_error
{ public: explicit ptree_error(const std::string& w) : std::runtime_error(w) {} };
class ptree_bad_path : public ptree_error
{ public: explicit ptree_bad_path(const std::string& w) : ptree_error(w) {} };
class ptree_bad_data : public ptree_error
{ public: explicit ptree_bad_data(const std::string& w) : ptree_error(w) {} };

namespace detail {
template <typename T> struct translator
{
    static bool from(const std::string& s, T& out)
    {
        std::istringstream iss(s);
        iss >> out;
        if(iss.fail() || iss.bad()) return false;
        iss >> std::ws;
        return iss.eof();
    }
};
template <> struct translator<std::string>
{
    static bool from(const std::string& s, std::string& out) { out = s; return true; }
};
template <> struct translator<bool>
{
    static bool from(const std::string& s, bool& out)
    {
        std::istringstream iss(s);
        iss >> out;
        if(iss.fail())
        {
            iss.clear();
            iss.str(s);
            iss.seekg(0);
            iss.setf(std::ios_base::boolalpha);
            iss >> out;
        }
        if(iss.fail() || iss.bad()) return false;
        iss >> std::ws;
        return iss.eof();
    }
};
} // namespace detail

class ptree
{
public:
    typedef std::string key_type;
    typedef std::string data_type;
    typedef std::pair<const std::string, ptree> value_type;
    typedef std::vector<value_type>::iterator iterator;
    typedef std::vector<value_type>::const_iterator const_iterator;

private:
    std::string data_;
    std::vector<value_type> kids_;

    const ptree* walk(const std::string& path) const
    {
        const ptree* cur = this;
        std::string::size_type start = 0;
        if(path.empty()) return cur;
        while(true)
        {
            std::string::size_type dot = path.find('.', start);
            std::string key = path.substr(start, dot == std::string::npos ? std::string::npos : dot - start);
            const ptree* next = nullptr;
            for(const auto& kv : cur->kids_)
                if(kv.first == key) { next = &kv.second; break; }
            if(!next) return nullptr;
            cur = next;
            if(dot == std::string::npos) break;
            start = dot + 1;
        }
        return cur;
    }

public:
    ptree() {}
    explicit ptree(const std::string& d) : data_(d) {}
    ptree(const ptree& o) : data_(o.data_)
    {
        kids_.reserve(o.kids_.size());
        for(const auto& kv : o.kids_) kids_.emplace_back(kv.first, kv.second);
    }
    ptree& operator=(const ptree& o)
    {
        if(this != &o)
        {
            data_ = o.data_;
            std::vector<value_type> tmp;
            tmp.reserve(o.kids_.size());
            for(const auto& kv : o.kids_) tmp.emplace_back(kv.first, kv.second);
            kids_.swap(tmp);
        }
        return *this;
    }

    std::size_t size() const { return kids_.size(); }
    bool empty() const { return kids_.empty(); }
    iterator begin() { return kids_.begin(); }
    iterator end() { return kids_.end(); }
    const_iterator begin() const { return kids_.begin(); }
    const_iterator end() const { return kids_.end(); }

    std::string& data() { return data_; }
    const std::string& data() const { return data_; }

    iterator push_back(const value_type& v) { kids_.emplace_back(v.first, v.second); return kids_.end() - 1; }
    ptree& add_child(const std::string& key, const ptree& child) { kids_.emplace_back(key, child); return kids_.back().second; }
    std::size_t count(const std::string& key) const
    {
        std::size_t n = 0;
        for(const auto& kv : kids_) if(kv.first == key) ++n;
        return n;
    }

    ptree& get_child(const std::string& path)
    {
        const ptree* p = walk(path);
        if(!p) throw ptree_bad_path("No such node (" + path + ")");
        return *const_cast<ptree*>(p);
    }
    const ptree& get_child(const std::string& path) const
    {
        const ptree* p = walk(path);
        if(!p) throw ptree_bad_path("No such node (" + path + ")");
        return *p;
    }
    const ptree& get_child(const std::string& path, const ptree& dflt) const
    {
        const ptree* p = walk(path);
        return p ? *p : dflt;
    }

    template <typename T> T get_value() const
    {
        T out;
        if(!detail::translator<T>::from(data_, out))
            throw ptree_bad_data("conversion of data to type failed: \"" + data_ + "\"");
        return out;
    }
    template <typename T> T get_value(const T& dflt) const
    {
        T out;
        return detail::translator<T>::from(data_, out) ? out : dflt;
    }
    template <typename T> T get(const std::string& path) const { return get_child(path).template get_value<T>(); }
    template <typename T> T get(const std::string& path, const T& dflt) const
    {
        const ptree* p = walk(path);
        if(!p) return dflt;
        T out;
        return detail::translator<T>::from(p->data_, out) ? out : dflt;
    }
    std::string get(const std::string& path, const char* dflt) const { return get<std::string>(path, std::string(dflt)); }

    template <typename T> ptree& put(const std::string& key, const T& v)
    {
        std::ostringstream oss;
        oss.precision(17);
        oss << v;
        for(auto& kv : kids_)
            if(kv.first == key) { kv.second.data_ = oss.str(); return kv.second; }
        kids_.emplace_back(key, ptree(oss.str()));
        return kids_.back().second;
    }
};

}} // namespace boost::property_tree

#endif
