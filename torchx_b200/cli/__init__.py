"""``torchx`` command line (reference torchx/cli/main.py:35-119).  Sub-commands on the single-box launch path:
run, status, log, describe, cancel, runopts, builtins, configure.  (tracker / list / delete operate on persistent
or remote state that local schedulers do not have.)"""
