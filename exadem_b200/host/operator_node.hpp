// operator_node.hpp — minimal stand-in for the Onika operator interface the reference plugs into
// (onika/scg/operator.h, operator_slot.h, operator_factory.h — external to /root/reference).
//
// Only the surface the hot-path operators use is provided: a named-slot OperatorNode with execute()/documentation(),
// ADD_SLOT(type, name, direction[, default][, DocString]), an OperatorNodeFactory registry keyed by the YAML operator
// name, and the ONIKA_AUTORUN_INIT static-registration macro.  In a real exaDEM build these operators are compiled
// against the genuine Onika headers instead: names, slots and execute() bodies stay as written here.
#pragma once
#include <any>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace onika_lite {

enum SlotDirection { INPUT, OUTPUT, INPUT_OUTPUT, PRIVATE };
struct DocString { std::string text; };
struct RequiredTag {};
struct OptionalTag {};
static constexpr RequiredTag REQUIRED{};
static constexpr OptionalTag OPTIONAL{};

class OperatorNode;

// A slot is a named, typed, shareable value (the reference connects slots of successive operators by name).
struct SlotBase {
  std::string name;
  SlotDirection dir;
  bool required = false;
  std::shared_ptr<std::any> value;  // shared between operators once connected
  virtual ~SlotBase() = default;
};

template <class T>
struct Slot : SlotBase {
  Slot(OperatorNode* owner, const char* n, SlotDirection d);
  Slot(OperatorNode* owner, const char* n, SlotDirection d, RequiredTag, DocString = {}) : Slot(owner, n, d) { required = true; }
  Slot(OperatorNode* owner, const char* n, SlotDirection d, OptionalTag, DocString = {}) : Slot(owner, n, d) {}
  Slot(OperatorNode* owner, const char* n, SlotDirection d, DocString) : Slot(owner, n, d) {}
  Slot(OperatorNode* owner, const char* n, SlotDirection d, T dflt, DocString = {}) : Slot(owner, n, d) { set(std::move(dflt)); }
  bool has_value() const { return value && value->has_value(); }
  void set(T v) { if (!value) value = std::make_shared<std::any>(); *value = std::move(v); }
  T& operator*() {
    if (!has_value()) throw std::runtime_error("slot '" + name + "' has no value");
    return *std::any_cast<T>(value.get());
  }
  T* operator->() { return &**this; }
};

class OperatorNode {
 public:
  virtual ~OperatorNode() = default;
  virtual void execute() = 0;
  virtual std::string documentation() const { return {}; }
  std::map<std::string, SlotBase*> slots;
  // connect every slot to the pool entry of the same name (what the YAML graph builder does in Onika)
  void bind(std::map<std::string, std::shared_ptr<std::any>>& pool) {
    for (auto& kv : slots) {
      auto it = pool.find(kv.first);
      if (it != pool.end()) {
        kv.second->value = it->second;
      } else {
        if (!kv.second->value) kv.second->value = std::make_shared<std::any>();
        pool[kv.first] = kv.second->value;
      }
    }
  }
  void check_required() const {
    for (auto& kv : slots)
      if (kv.second->required && !(kv.second->value && kv.second->value->has_value()))
        throw std::runtime_error("required slot '" + kv.first + "' is not set");
  }
};

template <class T>
Slot<T>::Slot(OperatorNode* owner, const char* n, SlotDirection d) {
  name = n;
  dir = d;
  owner->slots[name] = this;
}

class OperatorNodeFactory {
 public:
  using Maker = std::function<std::shared_ptr<OperatorNode>()>;
  static OperatorNodeFactory* instance() {
    static OperatorNodeFactory f;
    return &f;
  }
  void register_factory(const std::string& name, Maker m) { makers_[name] = std::move(m); }
  std::shared_ptr<OperatorNode> make_operator(const std::string& name) const {
    auto it = makers_.find(name);
    if (it == makers_.end()) throw std::runtime_error("unknown operator '" + name + "'");
    return it->second();
  }
  std::vector<std::string> names() const {
    std::vector<std::string> v;
    for (auto& kv : makers_) v.push_back(kv.first);
    return v;
  }

 private:
  std::map<std::string, Maker> makers_;
};

template <class Op>
std::shared_ptr<OperatorNode> make_simple_operator() { return std::make_shared<Op>(); }

}  // namespace onika_lite

#define ADD_SLOT(T, NAME, ...) ::onika_lite::Slot<T> NAME { this, #NAME, __VA_ARGS__ }
#define ONIKA_AUTORUN_INIT(tag)                      \
  static void edb_autorun_##tag();                   \
  namespace {                                        \
  struct edb_autorun_t_##tag {                       \
    edb_autorun_t_##tag() { edb_autorun_##tag(); }   \
  } edb_autorun_i_##tag;                             \
  }                                                  \
  static void edb_autorun_##tag()
