/* Empty shim: PGPLOT is absent; plotting is compiled out in the reference (PLOTTING 0). */
