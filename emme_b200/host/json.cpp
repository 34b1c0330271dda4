// json.cpp -- see json.hpp.  Independent implementation of the reference's JSON dialect.
#include "json.hpp"

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace emme {
namespace json {

const char* kind_name(Kind k) {
    switch (k) {
        case Kind::Null: return "ValueCategory::Null";
        case Kind::Int: return "ValueCategory::NumberInt";
        case Kind::Float: return "ValueCategory::NumberFloat";
        case Kind::Bool: return "ValueCategory::Boolean";
        case Kind::String: return "ValueCategory::String";
        case Kind::Array: return "ValueCategory::Array";
        case Kind::Object: return "ValueCategory::Object";
        case Kind::ComplexArray: return "ValueCategory::TypedArrayComplexDouble";
    }
    return "";
}

void Value::type_error(std::initializer_list<Kind> wanted) const {
    if (kind() == Kind::Null) throw std::runtime_error("Undefined Property");
    std::ostringstream oss;
    oss << (wanted.size() == 1 ? "Incorrect JSON type, requires: "
                               : "Incorrect JSON type, requires one of: ");
    for (Kind k : wanted) oss << kind_name(k) << ", ";
    oss << "actually: " << kind_name(kind());
    throw std::runtime_error(oss.str());
}

double Value::number() const {
    if (kind() == Kind::Float) return std::get<double>(v_);
    if (kind() == Kind::Int) return std::get<int>(v_);
    type_error({Kind::Float, Kind::Int});
}
const std::string& Value::as_string() const {
    if (kind() != Kind::String) type_error({Kind::String});
    return std::get<std::string>(v_);
}
bool Value::as_boolean() const {
    if (kind() != Kind::Bool) type_error({Kind::Bool});
    return std::get<bool>(v_);
}
const Value::Object& Value::as_object() const {
    if (kind() != Kind::Object) type_error({Kind::Object});
    return *std::get<std::shared_ptr<Object>>(v_);
}
Value::Object& Value::as_object() {
    if (kind() != Kind::Object) type_error({Kind::Object});
    return *std::get<std::shared_ptr<Object>>(v_);
}
const Value::Array& Value::as_array() const {
    if (kind() != Kind::Array) type_error({Kind::Array});
    return *std::get<std::shared_ptr<Array>>(v_);
}
Value::Array& Value::as_array() {
    if (kind() != Kind::Array) type_error({Kind::Array});
    return *std::get<std::shared_ptr<Array>>(v_);
}

const Value& Value::at(const std::string& key) const {
    if (kind() == Kind::Object) {
        const auto& o = as_object();
        auto it = o.find(key);
        if (it != o.end()) return it->second;
    }
    throw std::runtime_error("Failed to accessing key: " + key);
}
const Value& Value::at(std::size_t idx) const {
    if (kind() == Kind::Array && idx < as_array().size()) return as_array()[idx];
    throw std::runtime_error("Failed to accessing index: " + std::to_string(idx));
}
Value& Value::operator[](const std::string& key) { return as_object()[key]; }
bool Value::contains(const std::string& key) const {
    return kind() == Kind::Object && as_object().count(key) != 0;
}

Value Value::clone() const {
    switch (kind()) {
        case Kind::Array: {
            Value out = Value::array();
            for (const auto& e : as_array()) out.as_array().push_back(e.clone());
            return out;
        }
        case Kind::Object: {
            Value out = Value::object();
            for (const auto& [k, e] : as_object()) out.as_object().emplace(k, e.clone());
            return out;
        }
        case Kind::ComplexArray:
            return Value::complex_array(*std::get<std::shared_ptr<ComplexArray>>(v_));
        default:
            return *this;
    }
}

static void spaces(std::ostream& os, std::size_t n) {
    for (std::size_t i = 0; i < n; ++i) os << ' ';
}

std::string Value::dump() const {
    std::ostringstream oss;
    switch (kind()) {
        case Kind::Null: oss << "null"; break;
        case Kind::Bool: oss << (std::get<bool>(v_) ? "true" : "false"); break;
        case Kind::Int: oss << std::get<int>(v_); break;
        case Kind::Float: oss << std::get<double>(v_); break;
        case Kind::String: oss << '"' << std::get<std::string>(v_) << '"'; break;
        case Kind::Object: {
            oss << '{';
            bool first = true;
            for (const auto& [k, e] : as_object()) {
                if (!first) oss << ',';
                first = false;
                oss << '"' << k << "\":" << e.dump();
            }
            oss << '}';
            break;
        }
        case Kind::Array: {
            oss << '[';
            bool first = true;
            for (const auto& e : as_array()) {
                if (!first) oss << ',';
                first = false;
                oss << e.dump();
            }
            oss << ']';
            break;
        }
        case Kind::ComplexArray: {
            oss << '[';
            bool first = true;
            for (const auto& c : *std::get<std::shared_ptr<ComplexArray>>(v_)) {
                if (!first) oss << ',';
                first = false;
                oss << '[' << c.real() << ',' << c.imag() << ']';
            }
            oss << ']';
            break;
        }
    }
    return oss.str();
}

// Layout of the reference's pretty_print (src/JsonParser.cpp:255-328): 4-space indent,
// one member per line, "[ ]" / "{ }" for empty containers.
std::string Value::pretty_print(std::size_t indent) const {
    std::ostringstream oss;
    switch (kind()) {
        case Kind::Object: {
            const auto& o = as_object();
            if (o.empty()) return "{ }";
            oss << "{\n";
            std::size_t n = 0;
            for (const auto& [k, e] : o) {
                spaces(oss, indent + 4);
                oss << '"' << k << "\": " << e.pretty_print(indent + 4);
                oss << (++n < o.size() ? ",\n" : "\n");
            }
            spaces(oss, indent);
            oss << '}';
            break;
        }
        case Kind::Array: {
            const auto& a = as_array();
            if (a.empty()) return "[ ]";
            oss << "[\n";
            for (std::size_t i = 0; i < a.size(); ++i) {
                if (i) oss << ",\n";
                spaces(oss, indent + 4);
                oss << a[i].pretty_print(indent + 4);
            }
            oss << '\n';
            spaces(oss, indent);
            oss << ']';
            break;
        }
        case Kind::ComplexArray: {
            const auto& a = *std::get<std::shared_ptr<ComplexArray>>(v_);
            if (a.empty()) return "[ ]";
            oss << "[\n";
            for (std::size_t i = 0; i < a.size(); ++i) {
                if (i) oss << ",\n";
                spaces(oss, indent + 4);
                oss << '[' << a[i].real() << ", " << a[i].imag() << ']';
            }
            oss << '\n';
            spaces(oss, indent);
            oss << ']';
            break;
        }
        default:
            return dump();
    }
    return oss.str();
}

// ------------------------------------------------------------------ scanner / parser
namespace {

struct Scanner {
    std::istream& is;
    std::string file;
    int row = 1, col = 1;

    [[noreturn]] void lexical() const {
        std::ostringstream oss;
        oss << file << ':' << row << ':' << col << ": error: unrecognized token";
        throw std::runtime_error(oss.str());
    }
    [[noreturn]] void syntax(const std::string& content, int r, int c) const {
        std::ostringstream oss;
        oss << file << ':' << r << ':' << c << ": error: unexpected content '" << content << '\'';
        throw std::runtime_error(oss.str());
    }
    static bool ws(char c) { return c == '\t' || c == '\n' || c == '\r' || c == ' '; }
    static bool num_start(char c) { return (c >= '0' && c <= '9') || c == '-' || c == '+'; }
    static bool num_char(char c) { return c == '.' || c == 'E' || c == 'e' || num_start(c); }

    // next non-blank character, or 0 at end of input
    char skip() {
        char c;
        while (is.get(c)) {
            if (!ws(c)) return c;
            ++col;
            if (c == '\n') {
                ++row;
                col = 1;
            }
        }
        return 0;
    }
    void expect_word(char first, const char* rest) {
        for (const char* p = rest; *p; ++p) {
            char c;
            if (!is.get(c) || c != *p) lexical();
        }
        col += 1 + (int)std::string(rest).size();
        (void)first;
    }

    Value value(char c) {
        const int r0 = row, c0 = col;
        if (c == 0) syntax("", r0, 0);
        if (c == '{') {
            ++col;
            Value obj = Value::object();
            char d = skip();
            if (d == '}') {
                ++col;
                return obj;
            }
            for (;;) {
                if (d != '"') syntax(std::string(1, d), row, col);
                std::string key = string_body();
                char colon = skip();
                if (colon != ':') syntax(std::string(1, colon), row, col);
                ++col;
                obj.as_object().emplace(std::move(key), value(skip()));
                d = skip();
                ++col;
                if (d == '}') break;
                if (d != ',') syntax(std::string(1, d), row, col - 1);
                d = skip();
            }
            return obj;
        }
        if (c == '[') {
            ++col;
            Value arr = Value::array();
            char d = skip();
            if (d == ']') {
                ++col;
                return arr;
            }
            for (;;) {
                arr.as_array().push_back(value(d));
                d = skip();
                ++col;
                if (d == ']') break;
                if (d != ',') syntax(std::string(1, d), row, col - 1);
                d = skip();
            }
            return arr;
        }
        if (c == '"') return Value(string_body());
        if (num_start(c)) {
            std::string tok;
            bool is_float = false;
            do {
                tok.push_back(c);
                is_float |= c == '.';
                ++col;
            } while (is.get(c) && num_char(c));
            if (is) is.unget(); else is.clear();
            // the reference's rule: FLOAT iff the token contains '.', INTEGER goes through atoi
            if (is_float) return Value(std::atof(tok.c_str()));
            return Value(std::atoi(tok.c_str()));
        }
        if (c == 't') { expect_word(c, "rue"); return Value(true); }
        if (c == 'f') { expect_word(c, "alse"); return Value(false); }
        if (c == 'n') { expect_word(c, "ull"); return Value(); }
        if (c == '}' || c == ']' || c == ':' || c == ',') syntax(std::string(1, c), r0, c0);
        lexical();
    }
    // after the opening quote; no escape processing (like the reference)
    std::string string_body() {
        std::string s;
        char c;
        ++col;
        while (is.get(c) && c != '"') {
            s.push_back(c);
            ++col;
        }
        ++col;
        return s;
    }
};

}  // namespace

Value parse(std::istream& is, const std::string& filename) {
    Scanner sc{is, filename};
    char c = sc.skip();
    if (c == 0) sc.syntax("", sc.row, 0);
    Value v = sc.value(c);
    c = sc.skip();
    if (c != 0) sc.syntax(std::string(1, c), sc.row, sc.col);
    return v;
}

Value parse(const std::string& text) {
    std::istringstream ss(text);
    return parse(ss);
}

Value parse_file(const std::string& filename) {
    std::ifstream ifs(filename);
    if (!ifs) throw std::runtime_error("File " + filename + " not found");
    return parse(ifs, filename);
}

}  // namespace json
}  // namespace emme
