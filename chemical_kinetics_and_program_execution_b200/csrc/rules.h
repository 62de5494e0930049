// Host-side rule front end: turns a tape program into a p-independent table of flux rules.
//
// The reference runs a problem body (framework/problems.scm) symbolically with re-entrant
// continuations and recomputes every world probability on each right-hand side
// (framework/tape_multiverse.scm:669-838).  The *shape* of that world tree does not depend on p
// (only the pruning of zero-probability branches does, tm.scm:565), so this front end walks it
// once per (problem, cl_k) and records, for every leaf world that modified a tape, the sequence
// of unfold/choice steps that determine its probability plus the original/adjusted window
// indices that seed the window-extension forest on the device.
#pragma once

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace tapes {

enum Tape : int { PROGRAM_TAPE = 0, DATA_TAPE = 1 };  // data-tape? = #f / #t in the reference

// What a problem body may do (framework/gambit_macros.scm:99-125: tape-get, tape-set!, choose).
class Machine {
 public:
  virtual ~Machine() {}
  virtual int read(Tape t, int cell) = 0;
  virtual void write(Tape t, int cell, int symbol) = 0;
  // Picks one of n options with the given (unnormalised) weights; returns the option number.
  virtual int pick(const double* weights, int n) = 0;
};

typedef std::function<void(Machine&)> Body;

struct Problem {
  int alphabet = 0;
  Body body;
};

// One probability factor of a leaf world.
struct Step {
  enum Kind : uint8_t { UNFOLD = 0, CHOICE = 1 };
  uint8_t kind;
  uint8_t length;       // UNFOLD: effective correlation length L (tm.scm:502)
  uint32_t long_index;  // UNFOLD: index into marg_L   (tm.scm:556)
  uint32_t short_index; // UNFOLD: index into marg_L-1 (tm.scm:507-511)
  double prob;          // CHOICE: w_j / sum(w)        (gambit_macros.scm:75-86)
};

// A modified tape of a leaf world: base-A indices of the original and adjusted view
// (tm.scm:185-204, 1434-1438) and the visible length.
struct Seed {
  int length = 0;
  uint64_t orig = 0, adjusted = 0;
  bool changed() const { return orig != adjusted; }
};

struct FluxRule {
  std::vector<Step> steps;
  Seed tape[2];  // [PROGRAM_TAPE], [DATA_TAPE]
};

struct RuleTable {
  int alphabet = 0, cl_k = 0;
  std::vector<FluxRule> rules;  // only leaf worlds that changed at least one tape
  int64_t worlds_walked = 0;    // all world-tree nodes visited while enumerating
  int64_t leaf_worlds = 0;      // program completions (changed or not)
};

// Walks the full (unpruned) world tree of `problem` at correlation length cl_k.
RuleTable enumerate_rules(const Problem& problem, int cl_k);

// Several GPUs evaluate one problem by dealing its flux rules (the forests of different leaf
// worlds are independent, and a rule's steps carry everything its probability depends on).
// flux_rule_costs estimates the flux terms of every rule: one per length-k window overlapping a
// changed cell and per assignment of the window cells outside the view.  deal_flux_rules is the
// longest-processing-time assignment (descending cost to the least loaded part, ties to the lower
// index), the same on every rank.  rule_table_part keeps the rules of one part, in table order.
std::vector<double> flux_rule_costs(const RuleTable& table);
std::vector<int> deal_flux_rules(const std::vector<double>& costs, int n_parts);
RuleTable rule_table_part(const RuleTable& table, int part, int n_parts);

// Problem registry (framework/tapes_py_interface.scm:24-36).
const Problem* find_problem(const std::string& tag);
void register_problem(const std::string& tag, int alphabet, Body body);
void register_builtin_problems();  // canary + ex1..ex5 and variants
std::vector<std::string> registered_tags();

// Rule-table problems: a rewrite-rule set given as data instead of a compiled body.
struct RewriteRule {
  int tape;      // which tape is rewritten
  int span;      // 1..4 cells, cells 0..span-1
  int catalyst;  // -1, or the symbol cell 0 of the other tape must hold
  int pattern[4];
  int replacement[4];
  double rate;           // accept probability once the pattern matched
  double select_weight;  // weight of picking this rule
};
Body body_from_rewrite_rules(std::vector<RewriteRule> rules);

// Programs given as data: the decision tree of a problem body (what framework/problems.scm states
// with tape-get / tape-set! / choose, framework/gambit_macros.scm:99-125), so that a new problem
// needs no recompilation.  Node 0 is the entry; children have larger indices than their parent.
struct ProgramTree {
  enum Kind : int32_t { END = 0, READ = 1, WRITE = 2, PICK = 3 };
  std::vector<int32_t> kind;         // per node
  std::vector<int32_t> a, b, c;      // READ: tape, cell; WRITE: tape, cell, symbol; PICK: options
  std::vector<int32_t> first_child;  // READ: A children by symbol read; WRITE: 1; PICK: one per option
  std::vector<int32_t> first_weight; // PICK: offset of its (unnormalised) weights
  std::vector<int32_t> child;
  std::vector<double> weight;
};
// Throws std::runtime_error when the tree is malformed.
Body body_from_program(ProgramTree tree, int alphabet);

// The decision tree of any body (compiled, rewrite rules, or itself a tree): the body is run
// against every combination of outcomes of its reads and choices.  Cells a body has read or written
// are remembered, so a node exists only where an outcome is open.  Used by the Monte-Carlo
// simulator, which interprets the tree on the device.
ProgramTree trace_body(const Body& body, int alphabet, size_t max_nodes = (size_t)1 << 22);

}  // namespace tapes
