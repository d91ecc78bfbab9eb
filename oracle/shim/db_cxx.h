// Reference-harness stand-in for <db_cxx.h>: declarations only. The reference
// includes berkeleydb_store.hpp (logical_process.hpp:13) but never instantiates it
// (store_base.hpp:19 selects LevelDB), so the names only have to parse.
#pragma once
#include <cstddef>
#include <sys/types.h>
enum { DB_CREATE = 1, DB_INIT_LOCK = 2, DB_INIT_LOG = 4, DB_INIT_MPOOL = 8, DB_INIT_TXN = 16,
       DB_AUTO_COMMIT = 32, DB_DBT_MALLOC = 64, DB_SET_RANGE = 128, DB_PREV = 256, DB_NEXT = 512,
       DB_LAST = 1024, DB_FIRST = 2048, DB_RDONLY = 4096 };
enum DBTYPE { DB_BTREE = 1 };
class DbTxn {
 public:
  int commit(u_int32_t);
  int abort();
};
class DbEnv {
 public:
  explicit DbEnv(u_int32_t);
  int open(const char*, u_int32_t, int);
  int close(u_int32_t);
  int txn_begin(DbTxn*, DbTxn**, u_int32_t);
};
class Dbt {
 public:
  Dbt();
  Dbt(void*, u_int32_t);
  void* get_data() const;
  u_int32_t get_size() const;
  void set_flags(u_int32_t);
  void set_data(void*);
  void set_size(u_int32_t);
  void set_ulen(u_int32_t);
};
class Dbc {
 public:
  int get(Dbt*, Dbt*, u_int32_t);
  int close();
};
class Db {
 public:
  Db(DbEnv*, u_int32_t);
  int open(DbTxn*, const char*, const char*, DBTYPE, u_int32_t, int);
  int put(DbTxn*, Dbt*, Dbt*, u_int32_t);
  int get(DbTxn*, Dbt*, Dbt*, u_int32_t);
  int cursor(DbTxn*, Dbc**, u_int32_t);
  int close(u_int32_t);
};
