// Test scaffolding (see oracle/shim/Eigen/Core): the subset of boost::program_options that the reference's
// DecodingParams uses to parse its command line (ref: ASMC_SRC/SRC/DecodingParams.cpp:76-160, 164-275):
// long options "--name value" / "--name=value", bool switches, required and defaulted values, and one positional
// catch-all.  Option guessing is off, as the reference configures it.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

// the real boost headers leave ::isnan visible, and the reference calls it unqualified (DecodingParams.cpp:313)
using std::isnan;

namespace boost
{
namespace program_options
{
class error : public std::logic_error
{
public:
  explicit error(const std::string& what) : std::logic_error(what) {}
};

class value_semantic
{
public:
  virtual ~value_semantic() = default;
  virtual bool isSwitch() const = 0;
  virtual bool isRequired() const = 0;
  virtual bool isMulti() const = 0;
  virtual void parse(const std::string& name, const std::string& text) = 0;
  virtual void applyDefault() = 0;
  virtual bool hasDefault() const = 0;
};

namespace detail
{
template <class T> struct is_vector : std::false_type {
};
template <class T> struct is_vector<std::vector<T>> : std::true_type {
};
template <class T> T lexical(const std::string& name, const std::string& text)
{
  if constexpr (std::is_same_v<T, std::string>) {
    return text;
  } else {
    std::istringstream is(text);
    T v{};
    is >> v;
    if (is.fail() || (is.peek() != std::char_traits<char>::eof())) {
      if constexpr (std::is_floating_point_v<T>) {
        try {
          return static_cast<T>(std::stod(text));  // nan / inf
        } catch (...) {
        }
      }
      throw error("the argument ('" + text + "') for option '--" + name + "' is invalid");
    }
    return v;
  }
}
}  // namespace detail

template <class T> class typed_value : public value_semantic
{
  T* mTarget;
  bool mRequired = false, mHasDefault = false, mSwitch = false;
  T mDefault{};

public:
  T mStored{};
  explicit typed_value(T* target, const bool sw = false) : mTarget(target), mSwitch(sw) {}
  typed_value* required()
  {
    mRequired = true;
    return this;
  }
  typed_value* default_value(const T& v)
  {
    mDefault = v;
    mHasDefault = true;
    return this;
  }
  bool isSwitch() const override { return mSwitch; }
  bool isRequired() const override { return mRequired; }
  bool isMulti() const override { return detail::is_vector<T>::value; }
  bool hasDefault() const override { return mHasDefault; }
  void parse(const std::string& name, const std::string& text) override
  {
    if constexpr (detail::is_vector<T>::value) {
      mStored.push_back(detail::lexical<typename T::value_type>(name, text));
    } else if constexpr (std::is_same_v<T, bool>) {
      mStored = true;
    } else {
      mStored = detail::lexical<T>(name, text);
    }
    if (mTarget) {
      *mTarget = mStored;
    }
  }
  void applyDefault() override
  {
    mStored = mDefault;
    if (mTarget) {
      *mTarget = mDefault;
    }
  }
};

template <class T> typed_value<T>* value(T* target = nullptr)
{
  return new typed_value<T>(target);
}
inline typed_value<bool>* bool_switch(bool* target = nullptr)
{
  return (new typed_value<bool>(target, true))->default_value(false);
}

struct option_description {
  std::string name, help;
  std::shared_ptr<value_semantic> semantic;
};

class options_description;
class options_description_easy_init
{
  options_description* mOwner;

public:
  explicit options_description_easy_init(options_description* o) : mOwner(o) {}
  options_description_easy_init& operator()(const char* name, value_semantic* s, const char* help = "");
  options_description_easy_init& operator()(const char* name, const char* help);
};

class options_description
{
public:
  std::string caption;
  std::vector<option_description> options;
  options_description() = default;
  explicit options_description(std::string c) : caption(std::move(c)) {}
  options_description_easy_init add_options() { return options_description_easy_init(this); }
  options_description& add(const options_description& o)
  {
    options.insert(options.end(), o.options.begin(), o.options.end());
    return *this;
  }
  const option_description* find(const std::string& name) const
  {
    for (const auto& o : options) {
      if (o.name == name) {
        return &o;
      }
    }
    return nullptr;
  }
};
inline options_description_easy_init& options_description_easy_init::operator()(const char* name, value_semantic* s, const char* help)
{
  mOwner->options.push_back({name, help, std::shared_ptr<value_semantic>(s)});
  return *this;
}
inline options_description_easy_init& options_description_easy_init::operator()(const char* name, const char* help)
{
  mOwner->options.push_back({name, help, std::shared_ptr<value_semantic>(bool_switch())});
  return *this;
}
inline std::ostream& operator<<(std::ostream& os, const options_description& d)
{
  if (!d.caption.empty()) {
    os << d.caption << ":\n";
  }
  for (const auto& o : d.options) {
    os << "  --" << o.name << (o.semantic->isSwitch() ? "" : " arg") << "\t" << o.help << "\n";
  }
  return os;
}

class positional_options_description
{
public:
  std::string name;
  positional_options_description& add(const char* n, int)
  {
    name = n;
    return *this;
  }
};

namespace command_line_style
{
enum style_t { allow_guessing = 0x1000, default_style = 0x1fff };
}

struct variable_value {
  std::shared_ptr<value_semantic> semantic;
  template <class T> const T& as() const
  {
    auto* tv = dynamic_cast<typed_value<T>*>(semantic.get());
    if (!tv) {
      throw error("bad any cast");
    }
    return tv->mStored;
  }
};

struct parsed_options {
  const options_description* desc = nullptr;
  std::vector<std::pair<std::string, std::string>> values;  // (option, text)
};

class variables_map
{
public:
  std::map<std::string, variable_value> mValues;
  const options_description* mDesc = nullptr;
  std::size_t count(const std::string& name) const { return mValues.count(name); }
  const variable_value& operator[](const std::string& name) const { return mValues.at(name); }
};

class command_line_parser
{
  std::vector<std::string> mArgs;
  const options_description* mDesc = nullptr;
  std::string mPositional;

public:
  command_line_parser(int argc, char* argv[]) : mArgs(argv + (argc > 0 ? 1 : 0), argv + argc) {}
  command_line_parser& options(const options_description& d)
  {
    mDesc = &d;
    return *this;
  }
  command_line_parser& style(int) { return *this; }
  command_line_parser& positional(const positional_options_description& p)
  {
    mPositional = p.name;
    return *this;
  }
  parsed_options run() const
  {
    parsed_options out;
    out.desc = mDesc;
    for (std::size_t i = 0; i < mArgs.size(); ++i) {
      const std::string& a = mArgs[i];
      if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
        std::string name = a.substr(2), text;
        bool hasText = false;
        const std::size_t eq = name.find('=');
        if (eq != std::string::npos) {
          text = name.substr(eq + 1);
          name = name.substr(0, eq);
          hasText = true;
        }
        const option_description* o = mDesc->find(name);
        if (!o) {
          throw error("unrecognised option '--" + name + "'");
        }
        if (o->semantic->isSwitch()) {
          out.values.emplace_back(name, "");
        } else {
          if (!hasText) {
            if (i + 1 >= mArgs.size()) {
              throw error("the required argument for option '--" + name + "' is missing");
            }
            text = mArgs[++i];
          }
          out.values.emplace_back(name, text);
        }
      } else if (!mPositional.empty()) {
        out.values.emplace_back(mPositional, a);
      } else {
        throw error("too many positional options have been specified on the command line");
      }
    }
    return out;
  }
};

inline void store(const parsed_options& p, variables_map& vm)
{
  vm.mDesc = p.desc;
  for (const auto& kv : p.values) {
    const option_description* o = p.desc->find(kv.first);
    if (vm.mValues.count(kv.first) && !o->semantic->isMulti()) {
      throw error("option '--" + kv.first + "' cannot be specified more than once");
    }
    o->semantic->parse(kv.first, kv.second);
    vm.mValues[kv.first] = variable_value{o->semantic};
  }
  for (const auto& o : p.desc->options) {
    if (!vm.mValues.count(o.name) && o.semantic->hasDefault()) {
      o.semantic->applyDefault();
      vm.mValues[o.name] = variable_value{o.semantic};
    }
  }
}
inline void notify(variables_map& vm)
{
  if (!vm.mDesc) {
    return;
  }
  for (const auto& o : vm.mDesc->options) {
    if (o.semantic->isRequired() && !vm.mValues.count(o.name)) {
      throw error("the option '--" + o.name + "' is required but missing");
    }
  }
}
}  // namespace program_options
}  // namespace boost
