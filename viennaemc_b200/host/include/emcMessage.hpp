// Message collector of the drop-in host API.
// Interface mirrored: reference include/emcMessage.hpp:43-65 -- warnings accumulate
// until print(); an error prints everything and aborts the process.
#ifndef EMC_MESSAGE_HPP
#define EMC_MESSAGE_HPP

#include <cstdlib>
#include <iostream>
#include <string>

class emcMessage {
  std::string pending;
  bool fatal = false;
  emcMessage() = default;
  void push(const char *tag, const std::string &s, bool leadingBlankLine) {
    pending += std::string(leadingBlankLine ? "\n" : "") + "    " + tag + s + "\n";
  }

public:
  emcMessage(const emcMessage &) = delete;
  void operator=(const emcMessage &) = delete;
  static emcMessage &getInstance() {
    static emcMessage theOne;
    return theOne;
  }
  emcMessage &add(std::string s) { return addWarning(std::move(s)); }
  emcMessage &addWarning(std::string s) {
    push("WARNING: ", s, false);
    return *this;
  }
  emcMessage &addDebug(std::string s) {
    push("DEBUG: ", s, false);
    return *this;
  }
  emcMessage &addError(std::string s, bool shouldAbort = true) {
    push("ERROR: ", s, true);
    fatal = true;
    if (shouldAbort)
      print();
    return *this;
  }
  void print(std::ostream &out = std::cout) {
    out << pending;
    out.flush();
    pending.clear();
    if (fatal)
      std::abort();
  }
};

#endif
