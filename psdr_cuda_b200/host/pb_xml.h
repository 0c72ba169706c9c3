// psdr-b200 host layer: a small XML reader (elements + attributes, comments, declarations; text is ignored) — enough for
// the Mitsuba-style scene files SceneLoader accepts (src/scene/scene_loader.cpp uses pugixml for the same subset).
#pragma once
#include <cctype>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pbhost {

struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> children;

    bool has(const std::string &k) const { for (auto &a : attrs) if (a.first == k) return true; return false; }
    std::string attr(const std::string &k, const std::string &def = "") const { for (auto &a : attrs) if (a.first == k) return a.second; return def; }
    float attr_float(const std::string &k, float def) const { return has(k) ? std::stof(attr(k)) : def; }
    const XmlNode *child(const std::string &tag) const { for (auto &c : children) if (c->name == tag) return c.get(); return nullptr; }
    std::vector<const XmlNode *> all(const std::string &tag) const { std::vector<const XmlNode *> r; for (auto &c : children) if (c->name == tag) r.push_back(c.get()); return r; }
    // first child whose `name` attribute is one of the given names (find_child_by_name, scene_loader.cpp:66-77)
    const XmlNode *named(std::initializer_list<const char *> names) const {
        for (auto &c : children) { const std::string n = c->attr("name"); for (auto *k : names) if (n == k) return c.get(); }
        return nullptr;
    }
};

class XmlParser {
public:
    explicit XmlParser(const std::string &text) : s(text) {}
    std::unique_ptr<XmlNode> parse() {
        auto root = std::make_unique<XmlNode>();
        root->name = "#document";
        while (true) {
            skip_misc();
            if (i >= s.size()) break;
            if (s[i] != '<') fail();
            root->children.push_back(element());
        }
        return root;
    }

private:
    const std::string &s;
    size_t i = 0;
    [[noreturn]] void fail() const { throw std::runtime_error("XML parsing failed"); }
    void skip_ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    void skip_misc() {   // whitespace, text, comments, <? ?>, <! >
        while (i < s.size()) {
            if (s.compare(i, 4, "<!--") == 0) { size_t e = s.find("-->", i + 4); if (e == std::string::npos) fail(); i = e + 3; }
            else if (s.compare(i, 2, "<?") == 0) { size_t e = s.find("?>", i + 2); if (e == std::string::npos) fail(); i = e + 2; }
            else if (s.compare(i, 2, "<!") == 0) { size_t e = s.find('>', i + 2); if (e == std::string::npos) fail(); i = e + 1; }
            else if (s[i] != '<') ++i;
            else break;
        }
    }
    std::string ident() {
        size_t b = i;
        while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == ':' || s[i] == '.')) ++i;
        if (b == i) fail();
        return s.substr(b, i - b);
    }
    std::unique_ptr<XmlNode> element() {
        ++i;   // '<'
        auto n = std::make_unique<XmlNode>();
        n->name = ident();
        while (true) {
            skip_ws();
            if (i >= s.size()) fail();
            if (s[i] == '/') { if (i + 1 >= s.size() || s[i + 1] != '>') fail(); i += 2; return n; }
            if (s[i] == '>') { ++i; break; }
            std::string k = ident();
            skip_ws();
            if (i >= s.size() || s[i] != '=') fail();
            ++i; skip_ws();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) fail();
            const char q = s[i++];
            size_t e = s.find(q, i);
            if (e == std::string::npos) fail();
            n->attrs.emplace_back(k, s.substr(i, e - i));
            i = e + 1;
        }
        while (true) {
            skip_misc();
            if (i >= s.size()) fail();
            if (s.compare(i, 2, "</") == 0) {
                i += 2;
                if (ident() != n->name) fail();
                skip_ws();
                if (i >= s.size() || s[i] != '>') fail();
                ++i;
                return n;
            }
            n->children.push_back(element());
        }
    }
};

}  // namespace pbhost
