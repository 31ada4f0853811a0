/* ORACLE build shim: intentionally empty (the peripheral-library configuration is irrelevant on a host). */
