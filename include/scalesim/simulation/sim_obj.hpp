// Event / state base classes of the application surface (reference include/scalesim/simulation/sim_obj.hpp:18-135):
// sim_event_base (white/red colour + cancel flag), sim_event (accessor vtable, result_out = the committed "RE" line),
// sim_state, what_if<App>.
#ifndef SCALESIM_B200_SIMULATION_SIM_OBJ_HPP_
#define SCALESIM_B200_SIMULATION_SIM_OBJ_HPP_
#include <iostream>
#include <boost/serialization/serialization.hpp>
#include "scalesim/util/type.hpp"
namespace scalesim {

class sim_event_base {
 public:
  sim_event_base() : is_white_(true), is_cancel_(false) {}
  virtual ~sim_event_base() {}
  bool is_white() const { return is_white_; }
  bool is_red() const { return !is_white_; }
  void change_white() const { is_white_ = true; }
  void change_red() const { is_white_ = false; }
  bool is_cancel() const { return is_cancel_; }
  void change_cancel() const { is_cancel_ = true; }
 private:
  mutable bool is_white_;
  mutable bool is_cancel_;
  friend class boost::serialization::access;
  template <class Archive> void serialize(Archive& ar, unsigned int) { ar & is_white_; ar & is_cancel_; }
};

class sim_event {
 protected:
  sim_event() {}
  virtual ~sim_event() {}
 public:
  virtual long id() const { return -1; }
  virtual long source() const { return -1; }
  virtual long destination() const { return -1; }
  virtual bool end() const { return false; }
  virtual long receive_time() const { return -1; }
  virtual long send_time() const { return -1; }
  virtual int size() const { return -1; }
  virtual sim_event_base* base() const = 0;
  bool is_white() const { return base()->is_white(); }
  bool is_red() const { return base()->is_red(); }
  void change_white() const { base()->change_white(); }
  void change_red() const { base()->change_red(); }
  void change_cancel() const { base()->change_cancel(); }
  bool is_cancel() const { return base()->is_cancel(); }
  // one committed result line: RE,id,src,send,dst,recv[,START|,END]
  void result_out() const {
    std::cout << "RE," << id() << ',' << source() << ',' << send_time() << ',' << destination() << ','
              << receive_time();
    if (source() == -1) std::cout << ",START";
    else if (end()) std::cout << ",END";
    std::cout << std::endl;
  }
  bool operator==(const sim_event& r) const {
    return id() == r.id() && destination() == r.destination() && receive_time() == r.receive_time() &&
           send_time() == r.send_time();
  }
};

class sim_state {
 public:
  sim_state() {}
  virtual ~sim_state() {}
  virtual long id() const = 0;
  virtual int size() const { return -1; }
  void result_out() const {}
};

template <class App>
class what_if {
 public:
  what_if() : lp_id_(-1), time_(-1), delete_event_id_(-1) {}
  virtual ~what_if() {}
  what_if(long lp_id, long time) : lp_id_(lp_id), time_(time), delete_event_id_(-1) {}
  what_if(long lp_id, long time, const state<App>& s) : lp_id_(lp_id), time_(time), update_state_(s), delete_event_id_(-1) {}
  what_if(long lp_id, long time, const event<App>& e) : lp_id_(lp_id), time_(time), add_event_(e), delete_event_id_(-1) {}
  what_if(long lp_id, long time, long delete_event_id) : lp_id_(lp_id), time_(time), delete_event_id_(delete_event_id) {}
  long lp_id_;
  long time_;
  state<App> update_state_;
  event<App> add_event_;
  long delete_event_id_;
 private:
  friend class boost::serialization::access;
  template <class Archive> void serialize(Archive& ar, unsigned int) {
    ar & lp_id_; ar & time_; ar & update_state_; ar & add_event_; ar & delete_event_id_;
  }
};

}  // namespace scalesim
#endif
