// json.hpp -- reader/writer for EMME's input.json / output.json format (host side).
//
// Keeps the reference's file format AND its observable quirks ("config"):
//   * a number token is a FLOAT only if it contains '.', otherwise it is an INTEGER parsed
//     with atoi -- so "1e-6" reads as 1 and "10" as int 10 (reference
//     src/JsonParser.cpp:434-444, 558-571);
//   * a missing key throws std::runtime_error("Failed to accessing key: <k>")
//     (src/JsonParser.cpp:98-106), a wrong type "Incorrect JSON type, requires ..."
//     (include/JsonParser.h:63-76), reading an undefined value "Undefined Property";
//   * strings have no escape processing; objects are unordered;
//   * numbers print with the stream's default precision (6 significant digits),
//     pretty_print indents by 4 (src/JsonParser.cpp:203-328).
// The implementation is independent (std::variant tree, hand-written scanner).
#pragma once
#include <complex>
#include <iosfwd>
#include <memory>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

namespace emme {
namespace json {

enum class Kind { Null = 0, Int, Float, Bool, String, Array, Object, ComplexArray };
const char* kind_name(Kind k);  // "ValueCategory::NumberInt" ... (the reference's names)

class Value {
   public:
    using Object = std::unordered_map<std::string, Value>;
    using Array = std::vector<Value>;
    using ComplexArray = std::vector<std::complex<double>>;

    Value() = default;
    Value(int v) : v_(v) {}
    Value(double v) : v_(v) {}
    Value(bool v) : v_(v) {}
    Value(const char* s) : v_(std::string(s)) {}
    Value(std::string s) : v_(std::move(s)) {}
    static Value object() { Value x; x.v_ = std::make_shared<Object>(); return x; }
    static Value array(std::size_t n = 0) { Value x; x.v_ = std::make_shared<Array>(n); return x; }
    static Value complex_array(ComplexArray a) {
        Value x; x.v_ = std::make_shared<ComplexArray>(std::move(a)); return x;
    }

    Kind kind() const { return static_cast<Kind>(v_.index()); }
    bool is_object() const { return kind() == Kind::Object; }
    bool is_array() const { return kind() == Kind::Array; }
    bool is_number() const { return kind() == Kind::Int || kind() == Kind::Float; }
    bool is_string() const { return kind() == Kind::String; }
    bool is_boolean() const { return kind() == Kind::Bool; }

    // conversions; throw std::runtime_error with the reference's messages
    double number() const;
    operator double() const { return number(); }
    const std::string& as_string() const;
    bool as_boolean() const;
    const Object& as_object() const;
    Object& as_object();
    const Array& as_array() const;
    Array& as_array();

    const Value& at(const std::string& key) const;  // throws "Failed to accessing key: k"
    const Value& at(std::size_t idx) const;         // throws "Failed to accessing index: i"
    Value& operator[](const std::string& key);      // object: inserts a Null when absent
    Value& operator[](std::size_t idx) { return as_array()[idx]; }
    const Value& operator[](std::size_t idx) const { return as_array()[idx]; }
    bool contains(const std::string& key) const;

    Value clone() const;                    // deep copy (containers are shared otherwise)
    std::string dump() const;               // compact
    std::string pretty_print(std::size_t indent = 0) const;

   private:
    [[noreturn]] void type_error(std::initializer_list<Kind> wanted) const;
    std::variant<std::monostate, int, double, bool, std::string, std::shared_ptr<Array>,
                 std::shared_ptr<Object>, std::shared_ptr<ComplexArray>>
        v_;
};

Value parse(std::istream& is, const std::string& filename = {});
Value parse(const std::string& text);
Value parse_file(const std::string& filename);  // throws "File <name> not found"

}  // namespace json
}  // namespace emme
