/* Export macro of the uwspr blocks (same name and meaning as the reference's include/uwspr/api.h:25-31). */
#ifndef INCLUDED_UWSPR_API_H
#define INCLUDED_UWSPR_API_H

#include <gnuradio/attributes.h>

#ifdef gnuradio_uwspr_EXPORTS
#define UWSPR_API __GR_ATTR_EXPORT
#else
#define UWSPR_API __GR_ATTR_IMPORT
#endif

#endif
