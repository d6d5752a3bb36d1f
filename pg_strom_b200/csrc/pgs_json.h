/*
 * pgs_json.h - a small JSON value / parser / writer.
 *
 * The PostgreSQL glue hands plan fragments (expression trees, target lists)
 * to the planner half of this library as JSON text, the role nodeToString()
 * output would play inside a backend; results (rewritten target lists, the
 * partial-column catalogue) travel back the same way.  Only what that needs
 * is implemented: objects, arrays, strings, numbers, true/false/null.
 */
#ifndef PGS_JSON_H
#define PGS_JSON_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pgs {

struct Json;
typedef std::shared_ptr<Json> JsonPtr;

struct Json
{
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool        b = false;
    double      num = 0;
    std::string str;            /* String, and the raw text of a Number */
    std::vector<JsonPtr> arr;
    std::vector<std::pair<std::string, JsonPtr> > obj;  /* keeps order */

    static JsonPtr make(Kind k) { JsonPtr j(new Json); j->kind = k; return j; }
    static JsonPtr null() { return make(Null); }
    static JsonPtr boolean(bool v) { JsonPtr j = make(Bool); j->b = v; return j; }
    static JsonPtr number(double v)
    {
        JsonPtr j = make(Number);
        char buf[64];
        j->num = v;
        if (v == (double)(long long)v && v > -1e15 && v < 1e15)
            snprintf(buf, sizeof(buf), "%lld", (long long)v);
        else
            snprintf(buf, sizeof(buf), "%.17g", v);
        j->str = buf;
        return j;
    }
    static JsonPtr string(const std::string &s)
    { JsonPtr j = make(String); j->str = s; return j; }
    static JsonPtr array() { return make(Array); }
    static JsonPtr object() { return make(Object); }

    bool is_null() const { return kind == Null; }
    const Json *get(const std::string &key) const
    {
        for (size_t i = 0; i < obj.size(); i++)
            if (obj[i].first == key)
                return obj[i].second.get();
        return NULL;
    }
    JsonPtr getp(const std::string &key) const
    {
        for (size_t i = 0; i < obj.size(); i++)
            if (obj[i].first == key)
                return obj[i].second;
        return JsonPtr();
    }
    bool has(const std::string &key) const
    { const Json *j = get(key); return j && !j->is_null(); }
    std::string s(const std::string &key, const std::string &dflt = "") const
    {
        const Json *j = get(key);
        if (!j || j->is_null()) return dflt;
        return j->str;
    }
    long long i(const std::string &key, long long dflt = 0) const
    {
        const Json *j = get(key);
        if (!j || j->is_null()) return dflt;
        if (j->kind == Bool) return j->b ? 1 : 0;
        if (j->kind == String) return atoll(j->str.c_str());
        return (long long)j->num;
    }
    double d(const std::string &key, double dflt = 0) const
    {
        const Json *j = get(key);
        if (!j || j->is_null()) return dflt;
        if (j->kind == String) return atof(j->str.c_str());
        return j->num;
    }
    bool flag(const std::string &key, bool dflt = false) const
    {
        const Json *j = get(key);
        if (!j || j->is_null()) return dflt;
        if (j->kind == Bool) return j->b;
        if (j->kind == Number) return j->num != 0;
        return j->str == "true" || j->str == "t";
    }
    void set(const std::string &key, JsonPtr v)
    {
        for (size_t i = 0; i < obj.size(); i++)
            if (obj[i].first == key) { obj[i].second = v; return; }
        obj.push_back(std::make_pair(key, v));
    }
    void set(const std::string &key, const std::string &v) { set(key, string(v)); }
    void set(const std::string &key, const char *v) { set(key, string(v)); }
    void set(const std::string &key, long long v) { set(key, number((double)v)); }
    void set(const std::string &key, int v) { set(key, number((double)v)); }
    void setb(const std::string &key, bool v) { set(key, boolean(v)); }
    void push(JsonPtr v) { arr.push_back(v); }

    void write(std::string &out) const
    {
        switch (kind)
        {
            case Null: out += "null"; break;
            case Bool: out += (b ? "true" : "false"); break;
            case Number: out += str; break;
            case String: write_string(out, str); break;
            case Array:
                out += '[';
                for (size_t i = 0; i < arr.size(); i++)
                {
                    if (i) out += ',';
                    arr[i]->write(out);
                }
                out += ']';
                break;
            case Object:
                out += '{';
                for (size_t i = 0; i < obj.size(); i++)
                {
                    if (i) out += ',';
                    write_string(out, obj[i].first);
                    out += ':';
                    obj[i].second->write(out);
                }
                out += '}';
                break;
        }
    }
    std::string dump() const { std::string s; write(s); return s; }

    static void write_string(std::string &out, const std::string &s)
    {
        out += '"';
        for (size_t i = 0; i < s.size(); i++)
        {
            unsigned char c = (unsigned char)s[i];
            switch (c)
            {
                case '"': out += "\\\""; break;
                case '\\': out += "\\\\"; break;
                case '\n': out += "\\n"; break;
                case '\r': out += "\\r"; break;
                case '\t': out += "\\t"; break;
                default:
                    if (c < 0x20)
                    {
                        char buf[8];
                        snprintf(buf, sizeof(buf), "\\u%04x", c);
                        out += buf;
                    }
                    else
                        out += (char)c;
            }
        }
        out += '"';
    }
};

class JsonParser
{
    const char *p;
    const char *end;

    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    [[noreturn]] void fail(const char *msg)
    { throw std::runtime_error(std::string("JSON parse error: ") + msg); }

    JsonPtr value()
    {
        ws();
        if (p >= end) fail("unexpected end");
        switch (*p)
        {
            case '{': return object();
            case '[': return array();
            case '"': return Json::string(str());
            case 't':
                if (end - p >= 4 && !strncmp(p, "true", 4)) { p += 4; return Json::boolean(true); }
                fail("bad literal");
            case 'f':
                if (end - p >= 5 && !strncmp(p, "false", 5)) { p += 5; return Json::boolean(false); }
                fail("bad literal");
            case 'n':
                if (end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; return Json::null(); }
                fail("bad literal");
            default: return number();
        }
    }
    JsonPtr number()
    {
        const char *s = p;
        while (p < end && (strchr("+-0123456789.eE", *p) != NULL)) p++;
        if (p == s) fail("bad number");
        JsonPtr j = Json::make(Json::Number);
        j->str.assign(s, p - s);
        j->num = atof(j->str.c_str());
        return j;
    }
    std::string str()
    {
        std::string out;
        p++;    /* opening quote */
        while (p < end && *p != '"')
        {
            if (*p == '\\')
            {
                p++;
                if (p >= end) fail("bad escape");
                switch (*p)
                {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case 'r': out += '\r'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'u':
                    {
                        if (end - p < 5) fail("bad \\u");
                        unsigned cp = (unsigned)strtoul(std::string(p + 1, 4).c_str(), NULL, 16);
                        p += 4;
                        if (cp < 0x80) out += (char)cp;
                        else if (cp < 0x800)
                        { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
                        else
                        { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: out += *p;
                }
                p++;
            }
            else
                out += *p++;
        }
        if (p >= end) fail("unterminated string");
        p++;
        return out;
    }
    JsonPtr array()
    {
        JsonPtr j = Json::array();
        p++;
        ws();
        if (p < end && *p == ']') { p++; return j; }
        for (;;)
        {
            j->arr.push_back(value());
            ws();
            if (p < end && *p == ',') { p++; continue; }
            if (p < end && *p == ']') { p++; break; }
            fail("expected , or ]");
        }
        return j;
    }
    JsonPtr object()
    {
        JsonPtr j = Json::object();
        p++;
        ws();
        if (p < end && *p == '}') { p++; return j; }
        for (;;)
        {
            ws();
            if (p >= end || *p != '"') fail("expected key");
            std::string k = str();
            ws();
            if (p >= end || *p != ':') fail("expected :");
            p++;
            j->obj.push_back(std::make_pair(k, value()));
            ws();
            if (p < end && *p == ',') { p++; continue; }
            if (p < end && *p == '}') { p++; break; }
            fail("expected , or }");
        }
        return j;
    }
public:
    static JsonPtr parse(const std::string &text)
    {
        JsonParser ps;
        ps.p = text.data();
        ps.end = text.data() + text.size();
        JsonPtr j = ps.value();
        ps.ws();
        if (ps.p != ps.end) ps.fail("trailing characters");
        return j;
    }
};

}   /* namespace pgs */
#endif  /* PGS_JSON_H */
