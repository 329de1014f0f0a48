// toml_subset.hpp — the part of TOML that raytracer-server scene files use.
//
// The reference parses scenes with the `toml` crate into serde structs
// (src/scene.rs:292-348).  The scenes use: comments, [table] and [[array-of-tables]] headers,
// bare/quoted (dotted) keys, basic and literal strings, integers, floats (incl. inf/nan,
// underscores, exponents), booleans, (multi-line, nested) arrays with trailing commas and
// inline tables.  That subset is parsed here into a small value tree; anything else is a
// parse error with a line number, like LoadTomlError::Parse.
#pragma once

#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rtb {
namespace toml {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
    enum Kind { Table, Array, String, Integer, Float, Boolean } kind = Table;
    // insertion-ordered table
    std::vector<std::pair<std::string, ValuePtr>> items;
    std::vector<ValuePtr> elems;
    std::string str;
    long long integer = 0;
    double real = 0.0;
    bool boolean = false;
    bool inline_closed = false;   // inline tables / static arrays cannot be extended by headers
    bool array_of_tables = false;

    const Value* find(const std::string& key) const {
        for (auto& kv : items)
            if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    ValuePtr find_ptr(const std::string& key) {
        for (auto& kv : items)
            if (kv.first == key) return kv.second;
        return nullptr;
    }
    bool is_number() const { return kind == Integer || kind == Float; }
    // serde's f64 visitor accepts TOML integers as well
    double as_double() const { return kind == Integer ? (double)integer : real; }
};

class Parser {
public:
    explicit Parser(const std::string& text) : s_(text) {}

    ValuePtr parse() {
        root_ = std::make_shared<Value>();
        cur_ = root_;
        while (true) {
            skip_ws_comments_newlines();
            if (eof()) break;
            if (peek() == '[') parse_header();
            else {
                parse_keyval(cur_);
                skip_ws();
                if (!eof() && peek() == '#') skip_comment();
                if (!eof() && !consume_newline()) fail("expected end of line after value");
            }
        }
        return root_;
    }

private:
    const std::string& s_;
    size_t i_ = 0;
    int line_ = 1;
    ValuePtr root_, cur_;

    [[noreturn]] void fail(const std::string& msg) const {
        throw ParseError("TOML parse error at line " + std::to_string(line_) + ": " + msg);
    }
    bool eof() const { return i_ >= s_.size(); }
    char peek(size_t k = 0) const { return i_ + k < s_.size() ? s_[i_ + k] : '\0'; }
    void skip_ws() {
        while (!eof() && (peek() == ' ' || peek() == '\t')) ++i_;
    }
    void skip_comment() {
        while (!eof() && peek() != '\n') ++i_;
    }
    bool consume_newline() {
        if (peek() == '\n') { ++i_; ++line_; return true; }
        if (peek() == '\r' && peek(1) == '\n') { i_ += 2; ++line_; return true; }
        return false;
    }
    void skip_ws_comments_newlines() {
        while (!eof()) {
            skip_ws();
            if (peek() == '#') skip_comment();
            else if (!consume_newline()) break;
        }
    }
    static bool bare_key_char(char c) {
        return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '-';
    }

    std::string parse_basic_string() {  // opening quote already at peek()
        ++i_;
        std::string out;
        while (true) {
            if (eof() || peek() == '\n') fail("unterminated string");
            char c = s_[i_++];
            if (c == '"') break;
            if (c == '\\') {
                if (eof()) fail("unterminated escape");
                char e = s_[i_++];
                switch (e) {
                    case 'b': out += '\b'; break;
                    case 't': out += '\t'; break;
                    case 'n': out += '\n'; break;
                    case 'f': out += '\f'; break;
                    case 'r': out += '\r'; break;
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case 'u': case 'U': {
                        int n = e == 'u' ? 4 : 8;
                        unsigned long cp = 0;
                        for (int k = 0; k < n; ++k) {
                            char h = eof() ? '\0' : s_[i_++];
                            int d = (h >= '0' && h <= '9') ? h - '0' : (h >= 'a' && h <= 'f') ? h - 'a' + 10
                                    : (h >= 'A' && h <= 'F') ? h - 'A' + 10 : -1;
                            if (d < 0) fail("bad unicode escape");
                            cp = cp * 16 + (unsigned long)d;
                        }
                        append_utf8(out, cp);
                        break;
                    }
                    default: fail("unknown escape in string");
                }
            } else out += c;
        }
        return out;
    }
    static void append_utf8(std::string& out, unsigned long cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) {
            out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F));
        } else {
            out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F));
            out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F));
        }
    }
    std::string parse_literal_string() {
        ++i_;
        std::string out;
        while (true) {
            if (eof() || peek() == '\n') fail("unterminated literal string");
            char c = s_[i_++];
            if (c == '\'') break;
            out += c;
        }
        return out;
    }

    std::vector<std::string> parse_key_path() {
        std::vector<std::string> path;
        while (true) {
            skip_ws();
            if (peek() == '"') path.push_back(parse_basic_string());
            else if (peek() == '\'') path.push_back(parse_literal_string());
            else {
                size_t b = i_;
                while (!eof() && bare_key_char(peek())) ++i_;
                if (b == i_) fail("expected a key");
                path.push_back(s_.substr(b, i_ - b));
            }
            skip_ws();
            if (peek() == '.') { ++i_; continue; }
            break;
        }
        return path;
    }

    // descend (creating tables) along all but the last key
    ValuePtr descend(ValuePtr t, const std::vector<std::string>& path, size_t upto, bool for_header) {
        for (size_t k = 0; k < upto; ++k) {
            ValuePtr nxt = t->find_ptr(path[k]);
            if (!nxt) {
                nxt = std::make_shared<Value>();
                t->items.emplace_back(path[k], nxt);
            } else if (nxt->kind == Value::Array && nxt->array_of_tables && for_header) {
                if (nxt->elems.empty()) fail("empty array of tables");
                nxt = nxt->elems.back();
            } else if (nxt->kind != Value::Table || nxt->inline_closed) {
                fail("key '" + path[k] + "' is not a table");
            }
            t = nxt;
        }
        return t;
    }

    void parse_header() {
        ++i_;  // '['
        bool aot = false;
        if (peek() == '[') { aot = true; ++i_; }
        std::vector<std::string> path = parse_key_path();
        if (peek() != ']') fail("expected ']'");
        ++i_;
        if (aot) {
            if (peek() != ']') fail("expected ']]'");
            ++i_;
        }
        skip_ws();
        if (!eof() && peek() == '#') skip_comment();
        if (!eof() && !consume_newline()) fail("garbage after table header");
        ValuePtr parent = descend(root_, path, path.size() - 1, true);
        const std::string& last = path.back();
        ValuePtr ex = parent->find_ptr(last);
        if (aot) {
            if (!ex) {
                ex = std::make_shared<Value>();
                ex->kind = Value::Array;
                ex->array_of_tables = true;
                parent->items.emplace_back(last, ex);
            } else if (!(ex->kind == Value::Array && ex->array_of_tables)) {
                fail("'" + last + "' redefined as array of tables");
            }
            auto t = std::make_shared<Value>();
            ex->elems.push_back(t);
            cur_ = t;
        } else {
            if (!ex) {
                ex = std::make_shared<Value>();
                parent->items.emplace_back(last, ex);
            } else if (ex->kind != Value::Table || ex->inline_closed) {
                fail("duplicate table '" + last + "'");
            }
            cur_ = ex;
        }
    }

    void parse_keyval(ValuePtr table) {
        std::vector<std::string> path = parse_key_path();
        if (peek() != '=') fail("expected '=' after key");
        ++i_;
        skip_ws();
        ValuePtr v = parse_value();
        ValuePtr parent = descend(table, path, path.size() - 1, false);
        if (parent->find(path.back())) fail("duplicate key '" + path.back() + "'");
        parent->items.emplace_back(path.back(), v);
    }

    ValuePtr parse_value() {
        skip_ws();
        if (eof()) fail("expected a value");
        char c = peek();
        auto v = std::make_shared<Value>();
        if (c == '"') {
            if (peek(1) == '"' && peek(2) == '"') fail("multi-line strings are not supported");
            v->kind = Value::String;
            v->str = parse_basic_string();
            return v;
        }
        if (c == '\'') {
            if (peek(1) == '\'' && peek(2) == '\'') fail("multi-line strings are not supported");
            v->kind = Value::String;
            v->str = parse_literal_string();
            return v;
        }
        if (c == '[') {
            ++i_;
            v->kind = Value::Array;
            v->inline_closed = true;
            while (true) {
                skip_ws_comments_newlines();
                if (eof()) fail("unterminated array");
                if (peek() == ']') { ++i_; break; }
                v->elems.push_back(parse_value());
                skip_ws_comments_newlines();
                if (peek() == ',') { ++i_; continue; }
                if (peek() == ']') { ++i_; break; }
                fail("expected ',' or ']' in array");
            }
            return v;
        }
        if (c == '{') {
            ++i_;
            v->kind = Value::Table;
            skip_ws();
            if (peek() == '}') { ++i_; v->inline_closed = true; return v; }
            while (true) {
                skip_ws();
                parse_keyval(v);
                skip_ws();
                if (peek() == ',') { ++i_; continue; }
                if (peek() == '}') { ++i_; break; }
                fail("expected ',' or '}' in inline table");
            }
            v->inline_closed = true;
            return v;
        }
        // scalar token: up to a delimiter
        size_t b = i_;
        while (!eof()) {
            char d = peek();
            if (d == ',' || d == ']' || d == '}' || d == '#' || d == '\n' || d == '\r' || d == ' ' || d == '\t') break;
            ++i_;
        }
        std::string tok = s_.substr(b, i_ - b);
        if (tok.empty()) fail("expected a value");
        if (tok == "true" || tok == "false") {
            v->kind = Value::Boolean;
            v->boolean = tok == "true";
            return v;
        }
        parse_number(tok, *v);
        return v;
    }

    void parse_number(const std::string& tok, Value& v) {
        std::string body = tok;
        if (body == "inf" || body == "+inf") { v.kind = Value::Float; v.real = INFINITY; return; }
        if (body == "-inf") { v.kind = Value::Float; v.real = -INFINITY; return; }
        if (body == "nan" || body == "+nan" || body == "-nan") { v.kind = Value::Float; v.real = NAN; return; }
        std::string clean;
        for (size_t k = 0; k < body.size(); ++k) {
            char c = body[k];
            if (c == '_') {
                bool ok = k > 0 && k + 1 < body.size() && isdigit((unsigned char)body[k - 1]) &&
                          isdigit((unsigned char)body[k + 1]);
                if (!ok) fail("misplaced '_' in number '" + tok + "'");
                continue;
            }
            clean += c;
        }
        bool is_float = false, has_digit = false;
        for (size_t k = 0; k < clean.size(); ++k) {
            char c = clean[k];
            if (c >= '0' && c <= '9') { has_digit = true; continue; }
            if (c == '+' || c == '-') {
                if (k == 0 || clean[k - 1] == 'e' || clean[k - 1] == 'E') continue;
                fail("bad number '" + tok + "'");
            }
            if (c == '.') {
                bool ok = k > 0 && k + 1 < clean.size() && isdigit((unsigned char)clean[k - 1]) &&
                          isdigit((unsigned char)clean[k + 1]);
                if (!ok) fail("bad float '" + tok + "'");
                is_float = true;
                continue;
            }
            if (c == 'e' || c == 'E') { is_float = true; continue; }
            if ((c == 'x' || c == 'o' || c == 'b') && k == 1 && clean[0] == '0') {
                int base = c == 'x' ? 16 : (c == 'o' ? 8 : 2);
                char* end = nullptr;
                v.kind = Value::Integer;
                v.integer = std::strtoll(clean.c_str() + 2, &end, base);
                if (*end != 0 || clean.size() == 2) fail("bad integer '" + tok + "'");
                return;
            }
            fail("unsupported value '" + tok + "'");
        }
        if (!has_digit) fail("bad number '" + tok + "'");
        char* end = nullptr;
        if (is_float) {
            v.kind = Value::Float;
            v.real = std::strtod(clean.c_str(), &end);
        } else {
            v.kind = Value::Integer;
            v.integer = std::strtoll(clean.c_str(), &end, 10);
        }
        if (*end != 0) fail("bad number '" + tok + "'");
    }
};

inline ValuePtr parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace toml
}  // namespace rtb
