// mini_ptree.h — TEST INFRASTRUCTURE (oracle/_ref build only; never part of libctrfk.so).
//
// Stand-in for Boost.PropertyTree 1.65.1 (install.sh:59) as used by ctr_source/ctr_static.h:
// read_xml(stream, tree, trim_whitespace), get_child(key), iteration over (name, subtree) pairs,
// get<T>(key), get<T>(key, default), data().  Behaviour reproduced from Boost's documented
// XML mapping: element text -> data(), attributes -> child "<xmlattr>", comments -> child
// "<xmlcomment>", key lookup returns the first child of that name, a missing key or a value that
// does not parse completely throws (ptree_bad_path / ptree_bad_data).
#ifndef CTRB200_MINI_PTREE_H
#define CTRB200_MINI_PTREE_H
#include <istream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace boost { namespace property_tree {
struct ptree_error : std::runtime_error { explicit ptree_error(const std::string& w) : std::runtime_error(w) {} };
struct ptree_bad_path : ptree_error { explicit ptree_bad_path(const std::string& w) : ptree_error(w) {} };
struct ptree_bad_data : ptree_error { explicit ptree_bad_data(const std::string& w) : ptree_error(w) {} };

class ptree {
public:
    typedef std::pair<std::string, ptree> value_type;
    typedef std::vector<value_type>::iterator iterator;
    typedef std::vector<value_type>::const_iterator const_iterator;
    iterator begin() { return m_children.begin(); }
    iterator end() { return m_children.end(); }
    const_iterator begin() const { return m_children.begin(); }
    const_iterator end() const { return m_children.end(); }
    std::string& data() { return m_data; }
    const std::string& data() const { return m_data; }
    ptree& add_child(const std::string& k, const ptree& t) { m_children.push_back(value_type(k, t)); return m_children.back().second; }
    const ptree* find(const std::string& key) const
    {
        const size_t dot = key.find('.');
        const std::string head = key.substr(0, dot);
        for (const auto& c : m_children)
            if (c.first == head) return dot == std::string::npos ? &c.second : c.second.find(key.substr(dot + 1));
        return nullptr;
    }
    ptree& get_child(const std::string& key)
    {
        const ptree* p = find(key);
        if (!p) throw ptree_bad_path("No such node (" + key + ")");
        return const_cast<ptree&>(*p);
    }
    const ptree& get_child(const std::string& key) const
    {
        const ptree* p = find(key);
        if (!p) throw ptree_bad_path("No such node (" + key + ")");
        return *p;
    }
    template <class T> T get(const std::string& key) const
    {
        T v;
        const ptree& c = get_child(key);
        if (!parse(c.m_data, v)) throw ptree_bad_data("conversion of data to type failed (" + key + ")");
        return v;
    }
    template <class T> T get(const std::string& key, const T& dflt) const
    {
        const ptree* p = find(key);
        T v;
        return (p && parse(p->m_data, v)) ? v : dflt;
    }
private:
    template <class T> static bool parse(const std::string& s, T& v)
    {
        std::istringstream iss(s);
        iss >> v;
        if (!iss.eof()) iss >> std::ws;
        return !iss.fail() && !iss.bad() && iss.get() == std::char_traits<char>::eof();
    }
    static bool parse(const std::string& s, std::string& v) { v = s; return true; }
    std::string m_data;
    std::vector<value_type> m_children;
};

namespace xml_parser {
const int no_concat_text = 1, no_comments = 2, trim_whitespace = 4;
struct xml_parser_error : ptree_error { explicit xml_parser_error(const std::string& w) : ptree_error(w) {} };
namespace detail {
inline std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
inline std::string unescape(const std::string& s)
{
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] != '&') { o += s[i]; continue; }
        const size_t e = s.find(';', i);
        const std::string n = e == std::string::npos ? "" : s.substr(i + 1, e - i - 1);
        if (n == "lt") o += '<'; else if (n == "gt") o += '>'; else if (n == "amp") o += '&';
        else if (n == "quot") o += '"'; else if (n == "apos") o += '\''; else { o += s[i]; continue; }
        i = e;
    }
    return o;
}
struct Parser {
    const std::string& s;
    size_t i;
    int flags;
    void fail(const char* w) const { throw xml_parser_error(std::string("xml: ") + w); }
    bool starts(const char* t) const { return s.compare(i, std::char_traits<char>::length(t), t) == 0; }
    void content(ptree& node, const std::string& closing)
    {
        std::string text;
        for (;;) {
            if (i >= s.size()) { if (!closing.empty()) fail("unexpected end of data"); break; }
            if (s[i] != '<') { text += s[i++]; continue; }
            if (starts("<!--")) {
                const size_t e = s.find("-->", i + 4);
                if (e == std::string::npos) fail("unterminated comment");
                if (!(flags & no_comments)) { ptree c; c.data() = s.substr(i + 4, e - i - 4); node.add_child("<xmlcomment>", c); }
                i = e + 3;
            } else if (starts("<?")) {
                const size_t e = s.find("?>", i + 2);
                if (e == std::string::npos) fail("unterminated processing instruction");
                i = e + 2;
            } else if (starts("<![CDATA[")) {
                const size_t e = s.find("]]>", i + 9);
                if (e == std::string::npos) fail("unterminated CDATA");
                text += s.substr(i + 9, e - i - 9); i = e + 3;
            } else if (starts("<!")) {
                const size_t e = s.find('>', i);
                if (e == std::string::npos) fail("unterminated declaration");
                i = e + 1;
            } else if (starts("</")) {
                const size_t e = s.find('>', i);
                if (e == std::string::npos) fail("unterminated closing tag");
                if (trim(s.substr(i + 2, e - i - 2)) != closing) fail("mismatched closing tag");
                i = e + 1;
                break;
            } else {
                element(node);
            }
        }
        const std::string t = (flags & trim_whitespace) ? trim(text) : text;
        node.data() += unescape(t);
    }
    void element(ptree& parent)
    {
        ++i;   // '<'
        size_t b = i;
        while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != '>' && s[i] != '/') ++i;
        const std::string name = s.substr(b, i - b);
        if (name.empty()) fail("empty tag name");
        ptree node, attrs;
        bool has_attr = false, self_closed = false;
        for (;;) {
            while (i < s.size() && isspace((unsigned char)s[i])) ++i;
            if (i >= s.size()) fail("unterminated tag");
            if (s[i] == '>') { ++i; break; }
            if (s[i] == '/') { if (i + 1 >= s.size() || s[i + 1] != '>') fail("bad tag end"); i += 2; self_closed = true; break; }
            b = i;
            while (i < s.size() && s[i] != '=' && !isspace((unsigned char)s[i])) ++i;
            const std::string an = s.substr(b, i - b);
            while (i < s.size() && (isspace((unsigned char)s[i]) || s[i] == '=')) ++i;
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) fail("attribute value expected");
            const char q = s[i++];
            const size_t e = s.find(q, i);
            if (e == std::string::npos) fail("unterminated attribute");
            ptree av; av.data() = unescape(s.substr(i, e - i));
            attrs.add_child(an, av); has_attr = true;
            i = e + 1;
        }
        if (has_attr) node.add_child("<xmlattr>", attrs);
        if (!self_closed) content(node, name);
        parent.add_child(name, node);
    }
};
}  // namespace detail
template <class Tree> void read_xml(std::istream& in, Tree& t, int flags = 0)
{
    const std::string s((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    Tree root;
    detail::Parser p{s, 0, flags};
    p.content(root, "");
    root.data().clear();
    t = root;
}
}  // namespace xml_parser
using xml_parser::read_xml;
}}  // namespace boost::property_tree
#endif
