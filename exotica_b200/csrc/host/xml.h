// xml.h -- minimal XML reader (elements, attributes, text, comments, <?..?>, <!..>) for URDF/SRDF and
// EXOTica XML configs.  The reference uses tinyxml2 (exotica_core/include/exotica_core/loaders/xml_loader.h)
// and urdfdom; neither is available here, and this subset is all those formats need.
#pragma once
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace exotica
{
namespace xml
{
struct Element
{
    std::string name;
    std::vector<std::pair<std::string, std::string>> attributes;
    std::vector<std::shared_ptr<Element>> children;
    std::string text;  // concatenated character data directly inside this element

    const std::string* Attribute(const std::string& key) const
    {
        for (const auto& a : attributes)
            if (a.first == key) return &a.second;
        return nullptr;
    }
    std::string AttributeOr(const std::string& key, const std::string& fallback) const
    {
        const std::string* v = Attribute(key);
        return v ? *v : fallback;
    }
    std::shared_ptr<Element> FirstChild(const std::string& tag) const
    {
        for (const auto& c : children)
            if (c->name == tag) return c;
        return nullptr;
    }
    std::vector<std::shared_ptr<Element>> Children(const std::string& tag) const
    {
        std::vector<std::shared_ptr<Element>> out;
        for (const auto& c : children)
            if (c->name == tag) out.push_back(c);
        return out;
    }
};

class Parser
{
public:
    explicit Parser(const std::string& text) : s_(text), i_(0) {}

    std::shared_ptr<Element> ParseDocument()
    {
        SkipMisc();
        if (i_ >= s_.size() || s_[i_] != '<') throw std::runtime_error("XML: no root element");
        std::shared_ptr<Element> root = ParseElement();
        return root;
    }

private:
    const std::string& s_;
    size_t i_;

    static bool IsSpace(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
    static bool IsNameChar(char c) { return !(IsSpace(c) || c == '>' || c == '/' || c == '=' || c == '<'); }

    void SkipSpace()
    {
        while (i_ < s_.size() && IsSpace(s_[i_])) ++i_;
    }

    bool StartsWith(const char* lit) const { return s_.compare(i_, std::char_traits<char>::length(lit), lit) == 0; }

    void SkipUntil(const char* lit)
    {
        size_t p = s_.find(lit, i_);
        if (p == std::string::npos) throw std::runtime_error(std::string("XML: unterminated construct, expected ") + lit);
        i_ = p + std::char_traits<char>::length(lit);
    }

    // whitespace, comments, processing instructions, doctype
    void SkipMisc()
    {
        for (;;)
        {
            SkipSpace();
            if (StartsWith("<!--"))
                SkipUntil("-->");
            else if (StartsWith("<?"))
                SkipUntil("?>");
            else if (StartsWith("<!") && !StartsWith("<![CDATA["))
                SkipUntil(">");
            else
                return;
        }
    }

    static std::string Unescape(const std::string& in)
    {
        if (in.find('&') == std::string::npos) return in;
        std::string out;
        for (size_t k = 0; k < in.size(); ++k)
        {
            if (in[k] != '&')
            {
                out += in[k];
                continue;
            }
            static const char* ent[5] = {"&lt;", "&gt;", "&amp;", "&quot;", "&apos;"};
            static const char rep[5] = {'<', '>', '&', '"', '\''};
            bool hit = false;
            for (int e = 0; e < 5 && !hit; ++e)
            {
                const size_t n = std::char_traits<char>::length(ent[e]);
                if (in.compare(k, n, ent[e]) == 0)
                {
                    out += rep[e];
                    k += n - 1;
                    hit = true;
                }
            }
            if (!hit && k + 2 < in.size() && in[k + 1] == '#')
            {
                // numeric character reference &#NN; / &#xHH; (ASCII range: what attribute values use for tabs and newlines)
                const size_t semi = in.find(';', k);
                if (semi != std::string::npos && semi - k <= 8)
                {
                    const bool hex = in[k + 2] == 'x' || in[k + 2] == 'X';
                    const std::string digits = in.substr(k + (hex ? 3 : 2), semi - k - (hex ? 3 : 2));
                    char* end = nullptr;
                    const long code = strtol(digits.c_str(), &end, hex ? 16 : 10);
                    if (end && *end == 0 && !digits.empty() && code > 0 && code < 128)
                    {
                        out += (char)code;
                        k = semi;
                        hit = true;
                    }
                }
            }
            if (!hit) out += '&';
        }
        return out;
    }

    std::string ParseName()
    {
        const size_t b = i_;
        while (i_ < s_.size() && IsNameChar(s_[i_])) ++i_;
        if (i_ == b) throw std::runtime_error("XML: expected a name at offset " + std::to_string(b));
        return s_.substr(b, i_ - b);
    }

    std::shared_ptr<Element> ParseElement()
    {
        // s_[i_] == '<'
        ++i_;
        std::shared_ptr<Element> e = std::make_shared<Element>();
        e->name = ParseName();
        for (;;)
        {
            SkipSpace();
            if (i_ >= s_.size()) throw std::runtime_error("XML: unexpected end inside <" + e->name + ">");
            if (s_[i_] == '/')
            {
                if (i_ + 1 >= s_.size() || s_[i_ + 1] != '>') throw std::runtime_error("XML: malformed empty-element tag <" + e->name + ">");
                i_ += 2;
                return e;
            }
            if (s_[i_] == '>')
            {
                ++i_;
                break;
            }
            std::string key = ParseName();
            SkipSpace();
            if (i_ >= s_.size() || s_[i_] != '=') throw std::runtime_error("XML: attribute '" + key + "' without value in <" + e->name + ">");
            ++i_;
            SkipSpace();
            if (i_ >= s_.size() || (s_[i_] != '"' && s_[i_] != '\'')) throw std::runtime_error("XML: unquoted attribute '" + key + "'");
            const char quote = s_[i_++];
            const size_t b = i_;
            while (i_ < s_.size() && s_[i_] != quote) ++i_;
            if (i_ >= s_.size()) throw std::runtime_error("XML: unterminated attribute '" + key + "'");
            e->attributes.emplace_back(key, Unescape(s_.substr(b, i_ - b)));
            ++i_;
        }
        // content
        for (;;)
        {
            if (i_ >= s_.size()) throw std::runtime_error("XML: missing </" + e->name + ">");
            if (s_[i_] != '<')
            {
                const size_t b = i_;
                while (i_ < s_.size() && s_[i_] != '<') ++i_;
                e->text += Unescape(s_.substr(b, i_ - b));
                continue;
            }
            if (StartsWith("<!--"))
            {
                SkipUntil("-->");
                continue;
            }
            if (StartsWith("<![CDATA["))
            {
                const size_t b = i_ + 9;
                const size_t p = s_.find("]]>", b);
                if (p == std::string::npos) throw std::runtime_error("XML: unterminated CDATA");
                e->text += s_.substr(b, p - b);
                i_ = p + 3;
                continue;
            }
            if (StartsWith("<?"))
            {
                SkipUntil("?>");
                continue;
            }
            if (StartsWith("</"))
            {
                i_ += 2;
                const std::string closing = ParseName();
                if (closing != e->name) throw std::runtime_error("XML: </" + closing + "> closes <" + e->name + ">");
                SkipSpace();
                if (i_ >= s_.size() || s_[i_] != '>') throw std::runtime_error("XML: malformed closing tag </" + closing + ">");
                ++i_;
                return e;
            }
            e->children.push_back(ParseElement());
        }
    }
};

inline std::shared_ptr<Element> Parse(const std::string& text)
{
    Parser p(text);
    return p.ParseDocument();
}

inline std::string Trim(const std::string& s)
{
    size_t b = 0, e = s.size();
    while (b < e && (s[b] == ' ' || s[b] == '\t' || s[b] == '\n' || s[b] == '\r')) ++b;
    while (e > b && (s[e - 1] == ' ' || s[e - 1] == '\t' || s[e - 1] == '\n' || s[e - 1] == '\r')) --e;
    return s.substr(b, e - b);
}

}  // namespace xml
}  // namespace exotica
